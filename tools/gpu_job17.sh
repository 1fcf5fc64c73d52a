#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | grep -E "ours|ref32|passed|failed"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
SMILE_CONV_SPLIT=1 timeout 300 python tools/run_kernel.py conv8 7 2>&1 | tail -1
SMILE_CONV_SPLIT=1 timeout 300 python tools/conv_accuracy.py 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'train', d['train']['ms_per_step'], 'bf16', d['train_bf16']['ms_per_step'])"
