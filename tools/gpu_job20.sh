#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -3
timeout 600 python tools/train_breakdown.py > gpurun_out/r04e_train_breakdown_fp32.txt 2>&1
head -12 gpurun_out/r04e_train_breakdown_fp32.txt; grep "warp3d_bwd\[" gpurun_out/r04e_train_breakdown_fp32.txt | head
SMILE_WARP_BWD_PLAIN=1 timeout 600 python tools/train_breakdown.py 2>&1 | grep -E "^step|warp3d_bwd" | head -6
