#!/bin/bash
# hydro one-barrier schedule (fixed launch bounds) + PET L2 prefetch: parity tests, A/B bench, ncu capture
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_hydro_gpu.py tests/test_mhd_gpu.py tests/test_golden_gpu.py -m gpu -q > $O/r01r_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r01r_pytest.log
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), 'launches', d['gpu_launches'], d['clocks'])
" || tail -3 ${1%.json}.err; }
for v in 1 0; do
  EK_HYDRO_1B=$v timeout 300 python bench.py --workload blast3d --steps 20 --no-cpu-baseline > $O/r01r_blast3d_1b$v.json 2> $O/r01r_blast3d_1b$v.err
  line $O/r01r_blast3d_1b$v.json "blast3d 1b=$v"
done
EK_HYDRO_1B=1 timeout 300 python bench.py --workload rt3d --steps 10 --no-cpu-baseline > $O/r01r_rt3d_1b1.json 2> $O/r01r_rt3d_1b1.err
line $O/r01r_rt3d_1b1.json "rt3d 1b=1"
for v in 1 0; do
EK_MHD_PET_PF=$v timeout 300 python bench.py --workload ot3d --steps 10 --no-cpu-baseline > $O/r01r_ot3d_pf$v.json 2> $O/r01r_ot3d_pf$v.err
line $O/r01r_ot3d_pf$v.json "ot3d pf=$v"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hydro_step -s 3 -c 1 -f -o $O/r01r_hydro_1b \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r01r_ncu_h.log 2>&1
ls -la $O | tail -4
