#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -4 )
python tools/train_breakdown.py 2>&1 | grep -v Warn | head -12
