#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -2
timeout 600 python tools/train_breakdown.py 2>&1 | grep -E "^step|attn_bwd" | head -6
