#!/bin/bash
O=gpurun_out; mkdir -p $O
EK_MHD_STAGED=2 timeout 300 python -m pytest tests/test_mhd_gpu.py -x -q 2>&1 | tail -2
SWEEP='{"EK_MHD_STAGED": ["0", "2"]}' timeout 600 python profiles/mhd_sweep.py 512 2>$O/sweep4.err | tee $O/sweep4.jsonl
bash profiles/gpu_ncu_full.sh full_call call_kernel EK_MHD_STAGED=2
