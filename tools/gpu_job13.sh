#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_data.py -x -q -k "pipeline or prefetch" 2>&1 | tail -5 )
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/job13_bench.json 2> gpurun_out/job13_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/job13_bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'eager', d['eager'], 'launches', d['gpu_launches'])
    print('e2e', json.dumps(d['e2e'])[:700])
    print('roofline', d['roofline']['frac'], d['roofline']['launch_ms'], 'train', d['train']['ms_per_step'], 'train_bf16', d['train_bf16']['ms_per_step'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/job13_bench.err').read()[-2500:])
PY
