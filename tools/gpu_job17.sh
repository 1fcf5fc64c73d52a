#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "conv or upsample" 2>&1 | tail -2
SMILE_CONV_SPLIT_CHAINS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fp16_split" 2>&1 | tail -2
for CH in 0 1; do
echo "=== chains knob $CH"
SMILE_CONV_SPLIT_CHAINS=$CH timeout 300 python tools/conv_accuracy.py 2>&1 | head -1
SMILE_CONV_SPLIT_CHAINS=$CH timeout 300 python tools/run_kernel.py conv8 7 2>&1 | tail -1
SMILE_CONV_SPLIT_CHAINS=$CH timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | grep -E "ours vs|passed|failed"
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train --breakdown 2>gpurun_out/job17_breakdown.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'])"
grep -E "upsample2x" gpurun_out/job17_breakdown.txt | head -3
