#!/bin/bash
# round-2 closing evidence run: gated suite, sanitizers (split conv forced on at 32^3), ncu launch list, bench lines
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r04_pytest_gpu.txt; cat gpurun_out/r04_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -1
SMILE_CONV_SPLIT=2 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r04_memcheck.txt 2>&1; tail -2 gpurun_out/r04_memcheck.txt
SMILE_CONV_SPLIT=2 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r04_racecheck.txt 2>&1; tail -2 gpurun_out/r04_racecheck.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r04_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/job21_b.log 2>&1
python bench.py --breakdown > gpurun_out/r04_bench.json 2> gpurun_out/r04_bench_breakdown.txt; tail -c 300 gpurun_out/r04_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r04_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r04_bench_reference.json
python bench.py --config mindboggle --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r04_bench_mindboggle.json 2>/dev/null; head -c 200 gpurun_out/r04_bench_mindboggle.json
timeout 600 python tools/train_breakdown.py > gpurun_out/r04e_train_breakdown_fp32.txt 2>&1; head -3 gpurun_out/r04e_train_breakdown_fp32.txt
SMILE_TRAIN_DTYPE=bf16 timeout 600 python tools/train_breakdown.py > gpurun_out/r04f_train_breakdown_bf16.txt 2>&1; head -3 gpurun_out/r04f_train_breakdown_bf16.txt
SMILE_CONV_SPLIT=1 timeout 600 python tools/conv_compare.py > gpurun_out/r04_conv_compare.txt 2>&1; tail -2 gpurun_out/r04_conv_compare.txt
