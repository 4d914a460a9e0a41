#!/bin/bash
O=gpurun_out; mkdir -p $O
for w in implode2d rt3d implode3d ot2d; do
  timeout 600 python bench.py --workload $w --steps 20 --no-cpu-baseline > $O/r01_bench_$w.json 2> $O/r01_bench_$w.err
  python -c "
import json
d=json.loads(open('$O/r01_bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['config']['workload'][:70], 'ms/step', round(d['ms_per_step'],3), 'Gcell/s', round(d['value']/1e9,3), 'hbm frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,3))
" || tail -3 $O/r01_bench_$w.err
done
