#!/bin/bash
# hydro one-barrier schedule + PET with early loads: parity tests, A/B bench, ncu captures
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_hydro_gpu.py tests/test_mhd_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $O/r01q_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r01q_pytest.log
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), 'launches', d['gpu_launches'], d['clocks'])
" || tail -3 ${1%.json}.err; }
for v in 1 0; do
  EK_HYDRO_1B=$v timeout 300 python bench.py --workload blast3d --steps 20 --no-cpu-baseline > $O/r01q_blast3d_1b$v.json 2> $O/r01q_blast3d_1b$v.err
  line $O/r01q_blast3d_1b$v.json "blast3d 1b=$v"
  EK_HYDRO_1B=$v timeout 300 python bench.py --workload rt3d --steps 10 --no-cpu-baseline > $O/r01q_rt3d_1b$v.json 2> $O/r01q_rt3d_1b$v.err
  line $O/r01q_rt3d_1b$v.json "rt3d 1b=$v"
done
timeout 300 python bench.py --workload ot3d --steps 10 --no-cpu-baseline > $O/r01q_ot3d.json 2> $O/r01q_ot3d.err
line $O/r01q_ot3d.json "ot3d pet2"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hydro_step -s 3 -c 1 -f -o $O/r01q_hydro_1b \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r01q_ncu_h.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mhd_pet -s 3 -c 1 -f -o $O/r01q_mhd_pet \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/r01q_ncu_pet.log 2>&1
ls -la $O | tail -6
