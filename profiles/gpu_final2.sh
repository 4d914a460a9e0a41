#!/bin/bash
# after restoring the faster P+E+T kernel body: MHD GPU tier, bench, launch list, ncu --set full of the three step kernels
O=gpurun_out; mkdir -p $O; T=r01y
timeout 900 python -m pytest tests/test_mhd_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest_gpu.log
timeout 600 python bench.py --workload ot3d --steps 10 > $O/${T}_bench_ot3d.json 2> $O/${T}_bench_ot3d.err; cut -c1-330 $O/${T}_bench_ot3d.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_ot3d.csv \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_m.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mhd_pet|mhd_faces|mhd_update" -s 3 -c 3 -f -o $O/${T}_mhd \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_ms.log 2>&1
