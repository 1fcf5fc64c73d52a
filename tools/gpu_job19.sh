#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "upsample or end_to_end or cwm" 2>&1 | tail -2
timeout 600 python tools/run_upsample.py
