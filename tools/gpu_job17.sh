#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "conv" 2>&1 | tail -2
SMILE_CONV_SPLIT=1 timeout 600 python tools/conv_compare.py 2>&1 | grep -E "16->16|12->12|8->8 |total"
