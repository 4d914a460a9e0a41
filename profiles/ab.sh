#!/bin/bash
# A/B of two builds of libek_b200.so in ONE gpurun call (boxes and power caps differ by a few percent between calls):
#   bash profiles/ab.sh TAG "ENV_A" "ENV_B" [rounds] [bench args]
# e.g. bash profiles/ab.sh r02k "EK_B200_LIB=euler_kokkos_b200/lib/var/pet3b/libek_b200.so" "EK_MHD_PET_L2AHEAD=0" 3
O=gpurun_out; mkdir -p $O
TAG=$1; A=$2; B=$3; N=${4:-3}; shift 4
for r in $(seq 1 $N); do
  for V in A B; do
    if [ $V = A ]; then E="$A"; else E="$B"; fi
    env $E timeout 400 python bench.py --no-cpu-baseline --no-secondary --steps 10 "$@" 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V', '$E', 'ms/step %.3f' % d['ms_per_step'], 'clk', d['clocks']['sm_mhz'])" | tee -a $O/${TAG}_ab.txt
  done
done
