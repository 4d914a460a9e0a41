#!/bin/bash
O=gpurun_out; mkdir -p $O
EK_MHD_STAGED=1 timeout 300 python -m pytest tests/test_mhd_gpu.py -x -q 2>&1 | tail -4
SWEEP='{"EK_MHD_STAGED": ["0", "1"], "EK_MHD_YCHUNK": ["16"]}' timeout 600 python profiles/mhd_sweep.py 512 2>$O/sweep3.err | tee $O/sweep3.jsonl
bash profiles/gpu_ncu_full.sh full_staged staged EK_MHD_STAGED=1
