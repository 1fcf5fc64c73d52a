#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2 )
python bench.py --breakdown > gpurun_out/r04_bench.json 2> gpurun_out/r04_bench_breakdown.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'train', d['train']['ms_per_step'], 'bf16', d['train_bf16']['ms_per_step'], d['train_bf16']['value'])
PY
python bench.py --config mindboggle --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r04_bench_mindboggle.json 2>/dev/null; python -c "
import json; m=json.loads(open('gpurun_out/r04_bench_mindboggle.json').read().strip().splitlines()[-1]); print('mind', m['value'], m['ms_per_step'], m['e2e']['value'], m['roofline']['frac'], m['train']['ms_per_step'])"
