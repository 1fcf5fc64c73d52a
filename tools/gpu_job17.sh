#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "conv" 2>&1 | tail -3
for S in 1 2; do
echo "=== SMILE_CONV_SPLIT=$S"
SMILE_CONV_SPLIT=$S timeout 600 python tools/conv_compare.py 2>&1 | grep -E "160x192x160|80x96x80|total"
done
