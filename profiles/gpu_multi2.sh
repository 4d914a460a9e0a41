#!/bin/bash
N=2; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -4
for ov in 1 0; do
  EK_HALO_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 10 --warmup 3 --workload ot3d > $O/r01_bench_ot3d_n${N}_ov${ov}.json 2> $O/r01_bench_ot3d_n${N}_ov${ov}.err
  python -c "
import json
d=json.loads(open('$O/r01_bench_ot3d_n${N}_ov${ov}.json').read().strip().splitlines()[-1])
print('ot3d overlap=$ov n_gpus',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'Gcell/s',round(d['value']/1e9,3),'e2e',round(d['e2e']['value']/1e9,3))
" || tail -5 $O/r01_bench_ot3d_n${N}_ov${ov}.err
done
