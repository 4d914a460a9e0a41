#!/bin/bash
# A/B/.../N of several environments / builds of libek_b200.so in ONE gpurun call (boxes and power caps differ by a few
# percent between calls):   bash profiles/abn.sh TAG ROUNDS "ENV_1" "ENV_2" ...     (bench args: $ABN_ARGS)
# e.g. bash profiles/abn.sh r02x 2 "EK_MHD_RUI=0" "EK_MHD_RUI=1 EK_MHD_RUI_LAG=8" "EK_B200_LIB=euler_kokkos_b200/lib/var/uc4/libek_b200.so EK_MHD_RUI=1"
O=gpurun_out; mkdir -p $O
TAG=$1; N=$2; shift 2
for r in $(seq 1 $N); do
  i=0
  for E in "$@"; do
    i=$((i+1))
    env $E timeout 400 python bench.py --no-cpu-baseline --no-secondary --steps ${ABN_STEPS:-10} $ABN_ARGS 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('V$i', '$E', 'ms/step %.3f' % d['ms_per_step'], 'clk', d['clocks']['sm_mhz'])" | tee -a $O/${TAG}_ab.txt
  done
done
