#!/bin/bash
mkdir -p gpurun_out
SMILE_WGRAD_TC=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 -o gpurun_out/r04g_wgrad_tc -f python tools/run_wgrad_tc.py > gpurun_out/job19_ncu.log 2>&1
tail -2 gpurun_out/job19_ncu.log
