#!/bin/bash
# stage R: L2 prefetch of the compact state N planes ahead (y-chunked order)
O=gpurun_out; mkdir -p $O
line() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'Gcell/s', round(d['value']/1e9,3), d['clocks'])
" || tail -3 ${1%.json}.err; }
for v in 12 0 24 6; do
EK_MHD_R_PF=$v timeout 300 python bench.py --workload ot3d --steps 10 --no-cpu-baseline > $O/r01w_ot3d_rpf$v.json 2> $O/r01w_ot3d_rpf$v.err
line $O/r01w_ot3d_rpf$v.json "ot3d r_pf=$v"
done
EK_MHD_R_PF=12 timeout 300 python -m pytest tests/test_mhd_gpu.py -m gpu -q -k "one_step or variants" 2>&1 | tail -2
