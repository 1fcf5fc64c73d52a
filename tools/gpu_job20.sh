#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -x -q -k "tensor_cores or bf16 or train" 2>&1 | tail -3
SMILE_TRAIN_DTYPE=bf16 timeout 600 python tools/train_breakdown.py > gpurun_out/r04f_train_breakdown_bf16.txt 2>&1
head -12 gpurun_out/r04f_train_breakdown_bf16.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'train', d['train']['ms_per_step'], 'bf16', d['train_bf16']['ms_per_step'], d['train_bf16'].get('value'))"
