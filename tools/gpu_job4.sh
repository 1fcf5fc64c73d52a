#!/bin/bash
mkdir -p gpurun_out
for pr in 1 2; do
  echo "== promo $pr"; SMILE_TMA_PROMO=$pr python tools/run_kernel.py fused 10
  SMILE_TMA_PROMO=$pr ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fused_march2 -c 1 python tools/run_kernel.py fused 1 2>&1 | grep -E "dram__bytes|gpu__time"
done > gpurun_out/job4_promo.txt 2>&1
cat gpurun_out/job4_promo.txt
( python -m pytest tests/test_gpu_parity.py -x -q -k "fused or end_to_end or attention" 2>&1 | tail -4 )
python tools/run_kernel.py fused_l2 10
