#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "conv3d_oracle or end_to_end or encoder" 2>&1 | tail -2
timeout 300 python tools/run_kernel.py conv1 7 2>&1 | tail -1
