#!/bin/bash
# CFL reduction folded into the MHD update kernel: parity tests (incl. hydro schedule variants), A/B bench, launch list
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_mhd_gpu.py tests/test_hydro_gpu.py tests/test_golden_gpu.py -m gpu -q > $O/r01t_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $O/r01t_pytest.log
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), 'launches', d['gpu_launches'], d['clocks'])
" || tail -3 ${1%.json}.err; }
for v in 1 0; do
EK_MHD_DTFOLD=$v timeout 300 python bench.py --workload ot3d --steps 10 --no-cpu-baseline > $O/r01t_ot3d_fold$v.json 2> $O/r01t_ot3d_fold$v.err
line $O/r01t_ot3d_fold$v.json "ot3d dtfold=$v"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r01t_launches_ot3d.csv \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/r01t_ncu_m.log 2>&1
