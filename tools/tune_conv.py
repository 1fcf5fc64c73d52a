#!/usr/bin/env python
"""Sweep the register-tile configurations of the SIMT conv kernels over the layer shapes of one ModeT forward
(160x192x160) and print the time of each: input for the dispatch rules in conv_tma.cu / conv_tma_flat.cu."""
import os
import statistics
import sys

import torch

sys.path.insert(0, ".")
from smilecode_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
# (B, Cin, Cout, D, H, W, norm_on_load)
LAYERS = [(2, 4, 8, 160, 192, 160, False), (2, 8, 8, 160, 192, 160, True), (2, 8, 16, 80, 96, 80, False),
          (2, 16, 16, 80, 96, 80, True), (1, 6, 12, 80, 96, 80, False), (1, 12, 12, 80, 96, 80, True),
          (1, 12, 2, 80, 96, 80, True), (1, 12, 24, 40, 48, 40, False), (1, 24, 4, 40, 48, 40, True),
          (1, 48, 8, 20, 24, 20, True)]
CFGS = [None, "16:4", "16:2", "8:8", "8:4", "8:2", "4:8", "4:4", "4:2"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
os.environ["SMILE_CONV_TC"] = "0"


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3


for (B, cin, cout, D, H, W, norm) in LAYERS:
    x = torch.randn(B, cin, D, H, W, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.1
    b = torch.randn(cout, device=dev, generator=g) * 0.1
    st = torch.stack([x.double().sum((2, 3, 4)).flatten(), (x.double() ** 2).sum((2, 3, 4)).flatten()], 1).contiguous() if norm else None
    res = []
    for flat in (True, False):
        if flat:
            os.environ.pop("SMILE_CONV_NO_FLAT", None)
        else:
            os.environ["SMILE_CONV_NO_FLAT"] = "1"
        for cfg in CFGS:
            if cfg is None:
                os.environ.pop("SMILE_CONV_FORCE", None)
            else:
                if int(cfg.split(":")[0]) > max(4, 2 * cout):
                    continue
                os.environ["SMILE_CONV_FORCE"] = cfg
            try:
                t = timed(lambda: ops.conv3d(x, w, b, in_stats=st, want_stats=True))
            except Exception as e:  # configuration not instantiated for this width
                continue
            res.append((t, ("flat " if flat else "tiled ") + (cfg or "default")))
    res.sort()
    dflt = [r for r in res if r[1] == "flat default"][0][0]
    print(f"{cin:3d}->{cout:<3d} {D}x{H}x{W} B{B}: default {dflt:7.1f} us | best " + "  ".join(f"{n} {t:.1f}" for t, n in res[:4]))
