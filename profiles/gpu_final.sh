#!/bin/bash
# Round-end check of the committed state: full GPU tier, smoke, both bench arms, launch lists, ncu --set full of the step kernels
O=gpurun_out; mkdir -p $O; T=r01z
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $O/${T}_bench_blast3d.json 2> $O/${T}_bench_blast3d.err; echo "bench rc=$?"; cut -c1-400 $O/${T}_bench_blast3d.json
timeout 600 python bench.py --workload ot3d --steps 10 > $O/${T}_bench_ot3d.json 2> $O/${T}_bench_ot3d.err; cut -c1-400 $O/${T}_bench_ot3d.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; cut -c1-300 $O/${T}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_blast3d.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_b.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_ot3d.csv \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_m.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mhd_pet|mhd_faces|mhd_update" -s 3 -c 3 -f -o $O/${T}_mhd \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_ms.log 2>&1
ls -la $O | grep ${T}
