#!/bin/bash
# A/B of fused L1 kernel variants: parity tests + kernel timing
mkdir -p gpurun_out
for V in 8; do
  echo "=== SMILE_FUSED_V2=$V"
  SMILE_FUSED_V2=$V timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or end_to_end" 2>&1 | tail -2
done
for V in 0 8 0 8; do
  echo "=== SMILE_FUSED_V2=$V"
  SMILE_FUSED_V2=$V timeout 300 python tools/run_kernel.py fused 9 2>&1 | tail -1
  SMILE_FUSED_V2=$V timeout 300 python tools/run_kernel.py fused_l2 9 2>&1 | tail -1
done
SMILE_FUSED_V2=8 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deep: value', d['value'], 'roofline', d['roofline']['frac'], d['roofline']['launch_ms'])"
