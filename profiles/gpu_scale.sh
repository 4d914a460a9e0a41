#!/bin/bash
N=${1:-4}
O=gpurun_out; mkdir -p $O
for w in blast3d rt3d ot3d; do
  extra=""; [ $w != blast3d ] && extra="--size 256 --workload $w"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 $extra > $O/r01_scale_${w}_n$N.json 2> $O/r01_scale_${w}_n$N.err
  python -c "
import json
d=json.loads(open('$O/r01_scale_${w}_n$N.json').read().strip().splitlines()[-1])
print('$w N=$N', d['config']['workload'][:60], 'ms/step', round(d['ms_per_step'],3), 'Gcell/s', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3))
" || tail -5 $O/r01_scale_${w}_n$N.err
done
