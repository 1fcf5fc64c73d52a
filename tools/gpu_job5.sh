#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/job5_pytest.txt 2>&1
python bench.py --steps 20 --warmup 3 --comparators --breakdown > gpurun_out/job5_bench.json 2> gpurun_out/job5_bench_breakdown.txt
python bench.py --steps 10 --warmup 3 --config mindboggle > gpurun_out/job5_bench_mindboggle.json 2> gpurun_out/job5_bench_mindboggle.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/job5_bench_reference.json 2> gpurun_out/job5_bench_reference.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_march2 -c 1 -o gpurun_out/r03d_fused_v2 -f python tools/run_kernel.py fused 1 > gpurun_out/job5_ncu.log 2>&1
tail -3 gpurun_out/job5_pytest.txt; cat gpurun_out/job5_bench.json | cut -c1-1500
