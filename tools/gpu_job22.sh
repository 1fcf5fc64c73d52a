#!/bin/bash
# 8-GPU record of the end of round 2: LPBA config (incl. configs[3] train_bf16 block) and the Mindboggle-shape config (configs[4])
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r04_bench_n8.json 2> gpurun_out/r04_bench_n8.err
tail -c 1500 gpurun_out/r04_bench_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --config mindboggle --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r04_bench_mindboggle_n8.json 2> gpurun_out/r04_bench_mindboggle_n8.err
tail -c 1200 gpurun_out/r04_bench_mindboggle_n8.json
