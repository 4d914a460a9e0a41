#!/bin/bash
# usage: gpu_ncu_bench.sh <tag> <kernel-regex> <bench args...>   one --set full capture of one launch of a bench workload
O=gpurun_out; mkdir -p $O
TAG=$1; KR=$2; shift; shift
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s 3 -c 1 -f -o $O/$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $O/$TAG.log 2>&1
tail -2 $O/$TAG.log
