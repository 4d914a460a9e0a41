#!/bin/bash
# One GPU-box job: parity tests, bench lines, ncu launch lists and one --set full capture per solver.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash profiles/gpu_job.sh <tag>'
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python __graft_entry__.py > $O/${TAG}_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_blast3d.json 2> $O/${TAG}_bench_blast3d.err; echo "bench rc=$?"
cat $O/${TAG}_bench_blast3d.json
timeout 600 python bench.py --workload ot3d --steps 10 > $O/${TAG}_bench_ot3d.json 2> $O/${TAG}_bench_ot3d.err
cat $O/${TAG}_bench_ot3d.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_blast3d.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_b.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_ot3d.csv \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_m.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hydro_step -s 3 -c 1 -f -o $O/${TAG}_hydro_step \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_hs.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mhd_ -s 12 -c 6 -f -o $O/${TAG}_mhd \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_ms.log 2>&1
ls -la $O | tail -20
