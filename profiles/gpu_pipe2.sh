#!/bin/bash
# slab-pipelined MHD step with stream priorities U > R > P+E+T
O=gpurun_out; mkdir -p $O
for m in 1 8; do
EK_MHD_PIPE=$m EK_MHD_PIPE_PRIO=0,1,2 timeout 80 python bench.py --workload ot3d --steps 4 --no-cpu-baseline > $O/r01x_ot3d_pipe${m}_prio.json 2> $O/r01x_ot3d_pipe${m}_prio.err
python -c "
import json
d=json.loads(open('$O/r01x_ot3d_pipe${m}_prio.json').read().strip().splitlines()[-1])
print('pipe=$m prio ms/step', round(d['ms_per_step'],3), 'Gcell/s', round(d['value']/1e9,3), 'finite', d['e2e']['finite'], d['clocks'])
" || tail -3 $O/r01x_ot3d_pipe${m}_prio.err
done
