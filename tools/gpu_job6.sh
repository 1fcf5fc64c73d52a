#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_parity.py -x -q -s -k "bf16 or pipeline or prepared" 2>&1 | tail -25 ) > gpurun_out/job6_pytest.txt 2>&1
tail -12 gpurun_out/job6_pytest.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --train-dtype bf16 > gpurun_out/job6_bench_bf16.json 2> gpurun_out/job6_bench_bf16.err
tail -c 1500 gpurun_out/job6_bench_bf16.json; tail -5 gpurun_out/job6_bench_bf16.err
