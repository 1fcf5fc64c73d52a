#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_breakdown.py > gpurun_out/r04e_train_breakdown_fp32.txt 2>&1; head -3 gpurun_out/r04e_train_breakdown_fp32.txt
SMILE_TRAIN_DTYPE=bf16 timeout 600 python tools/train_breakdown.py > gpurun_out/r04f_train_breakdown_bf16.txt 2>&1; head -3 gpurun_out/r04f_train_breakdown_bf16.txt
python bench.py --breakdown > gpurun_out/r04_bench.json 2> gpurun_out/r04_bench_breakdown.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'train', d['train']['ms_per_step'], d['train']['value'], 'bf16', d['train_bf16']['ms_per_step'], d['train_bf16']['value'])
PY
