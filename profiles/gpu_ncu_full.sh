#!/bin/bash
# usage: gpu_ncu_full.sh <tag> <kernel-regex> [ENV=VAL ...]   one --set full capture of one launch at 512^3
O=gpurun_out; mkdir -p $O
TAG=$1; KR=$2; shift; shift
env "$@" SWEEP='{"X": ["1"]}' timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s 3 -c 1 -f -o $O/$TAG python profiles/mhd_sweep.py 512 > $O/$TAG.log 2>&1
tail -3 $O/$TAG.log
