#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r04_train_launches.csv python tools/train_breakdown.py > gpurun_out/job19.log 2>&1
tail -2 gpurun_out/job19.log
