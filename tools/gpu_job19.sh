#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_march_kernel -c 1 -o gpurun_out/r04b_conv_split -f python tools/run_kernel.py conv8 1 > gpurun_out/job19_ncu.log 2>&1
tail -2 gpurun_out/job19_ncu.log
