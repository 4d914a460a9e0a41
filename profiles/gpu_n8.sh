#!/bin/bash
# 8 GPUs of one box: MHD Orszag-Tang weak scaling (512^3 per GPU) and configs[3] (Rayleigh-Taylor 1024^3 as 8 z-slabs)
O=gpurun_out; mkdir -p $O
free -g > $O/r01_n8_free.txt; nproc >> $O/r01_n8_free.txt
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
echo "host MemAvailable ${avail} GB"
if [ "$avail" -lt 220 ]; then echo 'not enough host memory for 8 x (2 x 8.9 GB pinned)'; exit 1; fi
run() { # name, extra args
  local name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/r01_n8_$name.json 2> $O/r01_n8_$name.err
  python -c "
import json
d=json.loads(open('$O/r01_n8_$name.json').read().strip().splitlines()[-1])
print('$name', d['config']['workload'][:80], 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'Gcell/s', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), d['config'].get('decomposition'))
" || tail -5 $O/r01_n8_$name.err
}
run ot3d --workload ot3d
run rt3d_1024 --workload rt3d --shape 1024,1024,128
nvidia-smi topo -m > $O/r01_n8_topo.txt 2>&1
