#!/bin/bash
# round-2 first GPU pass: gated suite, full-size bisect, sanitizers, comparators
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/job1_gpu.txt
( time python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 ) > gpurun_out/job1_pytest.txt 2>&1
python tools/parity_bisect.py 160x192x160 8,4,2,1,1 > gpurun_out/job1_bisect_lpba.txt 2>&1
python tools/parity_bisect.py 160x192x224 8,4,2,1,1 > gpurun_out/job1_bisect_mb8.txt 2>&1
python tools/parity_bisect.py 160x192x224 6,6,6,1,1 > gpurun_out/job1_bisect_mb6.txt 2>&1
python tools/comparators.py --json gpurun_out/job1_comparators.json > gpurun_out/job1_comparators.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/job1_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/job1_racecheck.txt 2>&1
tail -3 gpurun_out/job1_pytest.txt
