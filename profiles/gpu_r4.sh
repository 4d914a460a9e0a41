#!/bin/bash
# hydro folded-halo-rows layout (28x14 tile, 16 warps): parity tests, A/B bench, ncu capture
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_hydro_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py -m gpu -q > $O/r01s_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r01s_pytest.log
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), 'launches', d['gpu_launches'], d['clocks'])
" || tail -3 ${1%.json}.err; }
for v in 1 0; do
  EK_HYDRO_FOLD=$v timeout 300 python bench.py --workload blast3d --steps 20 --no-cpu-baseline > $O/r01s_blast3d_fold$v.json 2> $O/r01s_blast3d_fold$v.err
  line $O/r01s_blast3d_fold$v.json "blast3d fold=$v"
done
EK_HYDRO_FOLD=1 timeout 300 python bench.py --workload rt3d --steps 10 --no-cpu-baseline > $O/r01s_rt3d_fold1.json 2> $O/r01s_rt3d_fold1.err
line $O/r01s_rt3d_fold1.json "rt3d fold=1"
EK_HYDRO_FOLD=1 timeout 300 python bench.py --workload implode3d --steps 20 --no-cpu-baseline > $O/r01s_implode3d_fold1.json 2> $O/r01s_implode3d_fold1.err
line $O/r01s_implode3d_fold1.json "implode3d fold=1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hydro_step -s 3 -c 1 -f -o $O/r01s_hydro_fold \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r01s_ncu_h.log 2>&1
ls -la $O | tail -4
