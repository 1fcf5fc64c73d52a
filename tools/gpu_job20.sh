#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_breakdown.py > gpurun_out/r04c_train_breakdown_fp32.txt 2>&1
head -45 gpurun_out/r04c_train_breakdown_fp32.txt
