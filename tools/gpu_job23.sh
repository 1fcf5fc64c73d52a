#!/bin/bash
# last check of the committed state: gated suite, smoke, default bench line
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2 )
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'clocks', d['clocks'])
print('train', d['train']['ms_per_step'], 'bf16', d['train_bf16']['ms_per_step'])
PY
