#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_compare.py 2>&1 | grep -v Warn > gpurun_out/job8_conv_compare.txt; tail -18 gpurun_out/job8_conv_compare.txt
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -x -q 2>&1 | tail -3 )
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --train-dtype bf16 > gpurun_out/job8_bench_bf16.json 2> gpurun_out/job8_bench_bf16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/job8_bench_bf16.json').read().strip().splitlines()[-1])
print('fp32 fwd', d['value'], 'bf16 fwd', json.dumps(d['bf16_forward']), 'train', json.dumps(d['train'])[:300])
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --train-dtype bf16 --train-batch 8 > gpurun_out/job8_bench_bf16_b8.json 2> gpurun_out/job8_bench_bf16_b8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/job8_bench_bf16_b8.json').read().strip().splitlines()[-1])
    print('B8 train', json.dumps(d['train'])[:400])
except Exception as e:
    print('B8 failed', e); print(open('gpurun_out/job8_bench_bf16_b8.err').read()[-1500:])
PY
