#!/bin/bash
# PET (fused prims+elec+trace) kernel: parity tests, bench with/without, launch list, one full ncu capture.
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_mhd_gpu.py tests/test_golden_gpu.py -m gpu -x -q > $O/r01p_pytest_mhd.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r01p_pytest_mhd.log
for pet in 1 0; do
  EK_MHD_PET=$pet timeout 300 python bench.py --workload ot3d --steps 10 --no-cpu-baseline > $O/r01p_bench_ot3d_pet$pet.json 2> $O/r01p_bench_ot3d_pet$pet.err
  python -c "
import json
d=json.loads(open('$O/r01p_bench_ot3d_pet$pet.json').read().strip().splitlines()[-1])
print('pet=$pet ms/step', round(d['ms_per_step'],3), 'Gcell/s', round(d['value']/1e9,3), 'launches', d['gpu_launches'])
" || tail -3 $O/r01p_bench_ot3d_pet$pet.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r01p_launches_ot3d.csv \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/r01p_ncu_m.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mhd_pet -s 3 -c 1 -f -o $O/r01p_mhd_pet \
   python bench.py --workload ot3d --steps 3 --warmup 3 --no-cpu-baseline > $O/r01p_ncu_pet.log 2>&1
ls -la $O | tail -8
