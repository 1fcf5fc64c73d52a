#!/usr/bin/env python
"""Per-layer conv3d timing at the LPBA shape: fp32 path (SIMT / 3xTF32 tcgen05) vs bf16 tcgen05 (kind::f16).
    python tools/conv_compare.py
Every encoder / CWM layer of ModeT.forward, both volumes of a pair (B = 2 for the encoder, 1 for the CWM)."""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from smilecode_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
LAYERS = [(2, 4, 8, (160, 192, 160)), (2, 8, 8, (160, 192, 160)), (2, 8, 16, (80, 96, 80)), (2, 16, 16, (80, 96, 80)),
          (2, 16, 32, (40, 48, 40)), (2, 32, 32, (40, 48, 40)), (2, 32, 64, (20, 24, 20)), (2, 64, 64, (20, 24, 20)),
          (2, 64, 128, (10, 12, 10)), (2, 128, 128, (10, 12, 10)), (1, 6, 12, (80, 96, 80)), (1, 12, 12, (80, 96, 80)),
          (1, 12, 24, (40, 48, 40)), (1, 24, 24, (40, 48, 40)), (1, 24, 48, (20, 24, 20)), (1, 48, 48, (20, 24, 20))]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3


tot = {"fp32": 0.0, "bf16": 0.0, "best": 0.0}
with torch.no_grad():
    for B, cin, cout, shp in LAYERS:
        x = torch.randn(B, cin, *shp, device=dev, generator=g)
        w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.05)
        b = torch.randn(cout, device=dev, generator=g) * 0.1
        st = torch.stack([x.double().sum((2, 3, 4)).flatten(), (x.double() ** 2).sum((2, 3, 4)).flatten()], 1).contiguous()
        t32 = timeit(lambda: ops.conv3d(x, w, b, in_stats=st, want_stats=True))
        with ops.conv_precision("bf16"):
            t16 = timeit(lambda: ops.conv3d(x, w, b, in_stats=st, want_stats=True))
        gf = 2 * 27 * cin * cout * B * shp[0] * shp[1] * shp[2] / 1e9
        print(f"{cin:4d}->{cout:<4d} {shp[0]}x{shp[1]}x{shp[2]:<4d} B={B}  fp32 {t32:8.1f} us  bf16 {t16:8.1f} us  "
              f"({gf / t16 * 1e3:7.1f} TFLOP/s bf16)  ratio {t32 / t16:5.2f}")
        tot["fp32"] += t32
        tot["bf16"] += t16
        tot["best"] += min(t32, t16)
print(f"total fp32 {tot['fp32']:.0f} us, bf16 {tot['bf16']:.0f} us, per-layer best {tot['best']:.0f} us")
