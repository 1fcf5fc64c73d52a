#!/usr/bin/env python
"""Launch the tensor-core weight gradient at one layer shape (ncu target / timing).
    SMILE_WGRAD_TC=2 python tools/run_wgrad_tc.py [cin cout D H W B]"""
import sys, statistics
import torch
sys.path.insert(0, ".")
from smilecode_b200 import ops
a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [8, 8, 160, 192, 160, 2]
cin, cout, D, H, W, B = a
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn(B, cin, D, H, W, device="cuda", generator=g)
gy = torch.randn(B, cout, D, H, W, device="cuda", generator=g)
w = torch.zeros(cout, cin, 3, 3, 3, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with ops.conv_precision("bf16"):
    for _ in range(2):
        ops.conv3d_bwd(gy, x, w, need_x=False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s, e in ev:
        flush.zero_(); s.record(); ops.conv3d_bwd(gy, x, w, need_x=False); e.record()
    torch.cuda.synchronize()
print(f"wgrad {cin}->{cout} {D}x{H}x{W} B={B}: median {statistics.median(s.elapsed_time(e) for s, e in ev) * 1e3:.1f} us")
