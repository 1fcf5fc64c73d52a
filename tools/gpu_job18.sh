#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | grep -E "ours|ref32|passed|failed" )
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/job18_bench.json 2> gpurun_out/job18_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/job18_bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'launches', d['gpu_launches'])
    print('e2e', d['e2e']['value'])
    print('roofline', d['roofline']['frac'], d['roofline']['launch_ms'])
    print('train', d.get('train',{}).get('ms_per_step'), 'train_bf16', d.get('train_bf16',{}).get('ms_per_step'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/job18_bench.err').read()[-2500:])
PY
