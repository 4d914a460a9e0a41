#!/bin/bash
# per-kernel time + DRAM/L2 traffic of one MHD step (512^3) for the env knobs given as arguments
O=gpurun_out; mkdir -p $O
TAG=$1; shift
env "$@" SWEEP='{"X": ["1"]}' timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:mhd_ -s 20 -c 5 --csv --log-file $O/${TAG}.csv python profiles/mhd_sweep.py 512 > $O/${TAG}.log 2>&1
python - <<PY
import csv
rows=list(csv.DictReader(l for l in open("$O/${TAG}.csv") if l.startswith('"')))
by={}
for r in rows:
    by.setdefault((r["ID"],r["Kernel Name"][:40]),{})[r["Metric Name"]]=(r["Metric Value"],r["Metric Unit"])
for (i,k),m in by.items():
    print(i,k,{a.split('.')[0].replace('__','_')[-28:]:b for a,b in m.items()})
PY
