#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_mhd_gpu.py -x -q 2>&1 | tail -3
timeout 600 python profiles/mhd_sweep.py 512 > $O/sweep1_default.jsonl 2>$O/sweep1_default.err
cat $O/sweep1_default.jsonl
for v in rminb3 rminb5 rminb6; do
  SWEEP='{"EK_MHD_YCHUNK": ["32"], "EK_MHD_FUSE_R": ["1"]}' EK_B200_LIB=euler_kokkos_b200/lib/var/$v/libek_b200.so timeout 300 python profiles/mhd_sweep.py 512 2>$O/sweep1_$v.err | tee $O/sweep1_$v.jsonl
done
