#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_parity.py -x -q -k "fused or end_to_end or attention" 2>&1 | tail -5 ) > gpurun_out/job3_pytest.txt 2>&1
{
echo "== v2 fast (12 warps)"; python tools/run_kernel.py fused 10
echo "== v2 fast (8 warps)"; SMILE_FUSED_V2=2 python tools/run_kernel.py fused 10
echo "== v2 safe"; SMILE_RUN_LN=0 python tools/run_kernel.py fused 10
echo "== v1"; SMILE_FUSED_V1=1 python tools/run_kernel.py fused 10
} > gpurun_out/job3_timing.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_march2 -c 1 -o gpurun_out/r03c_fused_v2 -f python tools/run_kernel.py fused 1 > gpurun_out/job3_ncu.log 2>&1
cat gpurun_out/job3_timing.txt; tail -3 gpurun_out/job3_pytest.txt
