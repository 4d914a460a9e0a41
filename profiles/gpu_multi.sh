#!/bin/bash
# usage: gpurun --gpus N -- bash profiles/gpu_multi.sh N tag
N=${1:-2}; TAG=${2:-multi}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu.py tests/test_hydro_gpu.py -m gpu -x -q 2>&1 | tail -4
for ov in 1 0; do
  EK_HALO_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 20 --warmup 3 > $O/${TAG}_bench_n${N}_ov${ov}.json 2> $O/${TAG}_bench_n${N}_ov${ov}.err
  python -c "
import json,sys
d=json.loads(open('$O/${TAG}_bench_n${N}_ov${ov}.json').read().strip().splitlines()[-1])
print('overlap=$ov n_gpus',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'Gcell/s',round(d['value']/1e9,3),'e2e',round(d['e2e']['value']/1e9,3),'launches',d['gpu_launches'])
" || tail -5 $O/${TAG}_bench_n${N}_ov${ov}.err
done
timeout 300 python bench.py --steps 20 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2>/dev/null; python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n1.json').read().strip().splitlines()[-1]); print('n=1 ms/step',round(d['ms_per_step'],3),'Gcell/s',round(d['value']/1e9,3))"
