#!/bin/bash
# PET early-load variant A/B, update kernel at 5 CTAs/SM: MHD parity tests, bench, launch lists
O=gpurun_out; mkdir -p $O
EK_MHD_PET_EARLY=1 timeout 600 python -m pytest tests/test_mhd_gpu.py tests/test_golden_gpu.py -m gpu -q > $O/r01v_pytest_early.log 2>&1; echo "pytest(early) rc=$?"; tail -3 $O/r01v_pytest_early.log
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), 'launches', d['gpu_launches'], d['clocks'])
" || tail -3 ${1%.json}.err; }
for v in 1 0; do
EK_MHD_PET_EARLY=$v timeout 300 python bench.py --workload ot3d --steps 10 --no-cpu-baseline > $O/r01v_ot3d_early$v.json 2> $O/r01v_ot3d_early$v.err
line $O/r01v_ot3d_early$v.json "ot3d early=$v"
EK_MHD_PET_EARLY=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r01v_launches_ot3d_early$v.csv \
   python bench.py --workload ot3d --steps 2 --warmup 3 --no-cpu-baseline > $O/r01v_ncu_m$v.log 2>&1
done
