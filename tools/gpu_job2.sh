#!/bin/bash
# fused L1 kernel v2: parity + timing + one ncu capture
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_parity.py -x -q -k "fused or end_to_end or attention" 2>&1 | tail -15 ) > gpurun_out/job2_pytest.txt 2>&1
{
echo "== v2 fast (12 warps)"; python tools/run_kernel.py fused 10
echo "== v2 fast (8 warps)"; SMILE_FUSED_V2=2 python tools/run_kernel.py fused 10
echo "== v2 safe (ln given, forced)"; SMILE_FUSED_V2=1 python tools/run_kernel.py fused 10
echo "== v2 safe (no ln)"; SMILE_RUN_LN=0 python tools/run_kernel.py fused 10
echo "== v1"; SMILE_FUSED_V1=1 python tools/run_kernel.py fused 10
echo "== v2 L2 level"; python tools/run_kernel.py fused_l2 10
echo "== v1 L2 level"; SMILE_FUSED_V1=1 python tools/run_kernel.py fused_l2 10
} > gpurun_out/job2_timing.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_march2 -c 1 -o gpurun_out/r03b_fused_v2 -f python tools/run_kernel.py fused 1 > gpurun_out/job2_ncu.log 2>&1
( python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | tail -15 ) > gpurun_out/job2_fullsize.txt 2>&1
cat gpurun_out/job2_timing.txt; tail -3 gpurun_out/job2_pytest.txt
