#!/bin/bash
mkdir -p gpurun_out
for c in 148 592; do echo "== SMILE_MARCH_CTAS=$c"; SMILE_MARCH_CTAS=$c timeout 300 python tools/conv_compare.py 2>&1 | grep -v Warn | head -4; done | tee gpurun_out/job7_conv_compare.txt
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "bf16" 2>&1 | tail -3 )
