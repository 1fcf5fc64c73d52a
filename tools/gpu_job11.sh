#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "qkrpb or modetqkrpb" 2>&1 | tail -3 )
python tools/comparators.py --json gpurun_out/job11_comparators.json 2>&1 | grep -v Warn | tail -32
