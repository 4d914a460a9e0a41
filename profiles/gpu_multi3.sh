#!/bin/bash
# 2 GPUs: slabs == single domain with the fused P+E+T kernel, the CFL fold and the folded hydro tile
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q > $O/r01u_pytest_multigpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r01u_pytest_multigpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus 2 --steps 10 --warmup 3 --workload ot3d --no-cpu-baseline > $O/r01u_ot3d_n2.json 2> $O/r01u_ot3d_n2.err
python -c "
import json
d=json.loads(open('$O/r01u_ot3d_n2.json').read().strip().splitlines()[-1])
print('ot3d n_gpus',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'Gcell/s',round(d['value']/1e9,3),'e2e',round(d['e2e']['value']/1e9,3))
" || tail -5 $O/r01u_ot3d_n2.err
