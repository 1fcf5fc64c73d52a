#!/usr/bin/env python
"""Per-layer accuracy of the conv3d kernels against an fp64 convolution (torch on the GPU): rms and max error relative
to the rms of the output, for the kernel the current environment selects (SMILE_CONV_SPLIT=0 -> SIMT fp32 FMA,
default -> fp16-split tensor cores where eligible).  Also prints torch's own fp32 conv (cuDNN, TF32 off) for scale.

    python tools/conv_accuracy.py [cin cout D H W]
"""
import sys

import torch

sys.path.insert(0, ".")
from smilecode_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
a = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else [8, 8, 40, 96, 160]
cin, cout, shape = a[0], a[1], tuple(a[2:5])
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.randn(2, cin, *shape, device="cuda", generator=g)
w0 = torch.randn(cin, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5
b0 = torch.randn(cin, device="cuda", generator=g) * 0.1
b = torch.randn(cout, device="cuda", generator=g) * 0.1
with torch.no_grad():
    raw, st = ops.conv3d(x, w0, b0, want_stats=True)          # producer layer (gives realistic IN statistics)
    out, _ = ops.conv3d(raw, w, b, in_stats=st)
    r64 = raw.double()
    mean = r64.mean((2, 3, 4), keepdim=True)
    var = r64.var((2, 3, 4), unbiased=False, keepdim=True)
    act = torch.nn.functional.leaky_relu((r64 - mean) / torch.sqrt(var + 1e-5), 0.1)
    ref = torch.nn.functional.conv3d(act, w.double(), b.double(), padding=1)
    t32 = torch.nn.functional.conv3d(act.float(), w, b, padding=1)
    scale = float(ref.pow(2).mean().sqrt())
    for name, y in (("ours", out), ("torch fp32", t32)):
        e = (y.double() - ref)
        print(f"{name:11s} {cin}->{cout} {shape}: rms err / rms out {float(e.pow(2).mean().sqrt()) / scale:.3e}   "
              f"max err / rms out {float(e.abs().max()) / scale:.3e}   mean err / rms out {float(e.mean()) / scale:+.3e}")
