#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_mhd_gpu.py -x -q 2>&1 | tail -2
SWEEP='{"EK_MHD_YCHUNK": ["16", "32", "64"], "EK_MHD_ZGROUP": ["0", "1"]}' timeout 600 python profiles/mhd_sweep.py 512 2>$O/sweep2.err | tee $O/sweep2.jsonl
bash profiles/gpu_ncu_mhd.sh ncu_zg1 EK_MHD_YCHUNK=32 EK_MHD_ZGROUP=1
