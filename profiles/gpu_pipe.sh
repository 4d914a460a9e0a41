#!/bin/bash
# slab-pipelined MHD step (EK_MHD_PIPE): variants parity test + bench at two stage sizes
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_mhd_gpu.py -m gpu -q -k "variants or free_run or conserv" > $O/r01x_pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r01x_pytest_pipe.log
for m in 2 4; do
EK_MHD_PIPE=$m timeout 100 python bench.py --workload ot3d --steps 6 --no-cpu-baseline > $O/r01x_ot3d_pipe$m.json 2> $O/r01x_ot3d_pipe$m.err
python -c "
import json
d=json.loads(open('$O/r01x_ot3d_pipe$m.json').read().strip().splitlines()[-1])
print('pipe=$m ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), 'finite', d['e2e']['finite'], 't_sim', d['t_sim'], d['clocks'])
" || tail -3 $O/r01x_ot3d_pipe$m.err
done
