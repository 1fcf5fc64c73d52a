#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -2
timeout 600 python tools/train_breakdown.py 2>&1 | grep -E "^step|wgrad" | head -12
SMILE_WGRAD_ONE_ROW=1 timeout 600 python tools/train_breakdown.py 2>&1 | grep -E "^step|wgrad" | head -5
