#!/bin/bash
# 8-GPU pass: copy-bandwidth probe (1 rank alone, then 8 at once), bench at N=8 (e2e modes, train, train_bf16 configs[3])
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python tools/pcie_probe.py 2>&1 | grep PCIE > gpurun_out/job10_pcie.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py 2>&1 | grep PCIE >> gpurun_out/job10_pcie.txt
cat gpurun_out/job10_pcie.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/job10_bench_n8.json 2> gpurun_out/job10_bench_n8.err
tail -c 3000 gpurun_out/job10_bench_n8.json; tail -3 gpurun_out/job10_bench_n8.err
nvidia-smi topo -m > gpurun_out/job10_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/job10_topo.txt; free -g >> gpurun_out/job10_topo.txt
