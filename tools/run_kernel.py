#!/usr/bin/env python
"""Launch one hot-path kernel at the LPBA shape a few times (target for `ncu -k regex:...`), and
print its CUDA-event time with L2 flushed between launches.

    python tools/run_kernel.py fused|fused_l2|conv8|conv4|warp8|proj|encoder [reps]
"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from smilecode_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fused"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
S = (160, 192, 160)
S2 = (80, 96, 80)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def smooth_flow(shape, amp):
    c = torch.randn(1, 3, *[max(2, s // 16) for s in shape], device=dev, generator=g) * amp
    return torch.nn.functional.interpolate(c, size=shape, mode="trilinear", align_corners=True).contiguous()


if what in ("fused", "fused_l2", "fused_nomov"):
    shp = S2 if what == "fused_l2" else S
    q = torch.randn(1, *shp, 6, device=dev, generator=g)
    k = torch.randn(1, *shp, 6, device=dev, generator=g)
    rpb = torch.randn(1, 3, 3, 3, device=dev, generator=g) * 0.5
    flow = smooth_flow(shp, 2.0)
    mov = torch.rand(1, 1, *shp, device=dev, generator=g) if what == "fused" else None
    import os
    lnp = {}
    if os.environ.get("SMILE_RUN_LN", "1") == "1":      # q, k ~ N(0,1) per channel: |q|_2 <= 6 is NOT guaranteed for raw randn,
        q = torch.nn.functional.layer_norm(q, (6,))     # so make them real LayerNorm outputs (gamma = 1, beta = 0)
        k = torch.nn.functional.layer_norm(k, (6,))
        lnp = {"ln_gamma": torch.ones(6, device=dev), "ln_beta": torch.zeros(6, device=dev)}
    fn = lambda: ops.modet_fused(q, k, rpb, flow, mov, 1.0, 1.0 if what == "fused" else 2.0, **lnp)
    nbytes = (80 if what == "fused" else 72) * shp[0] * shp[1] * shp[2]
    if what == "fused_nomov":
        ops.modet_attention(q, k, rpb, 1, 1.0)
        import time
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a_, b_ in evs:
            flush.zero_(); a_.record(); ops.modet_attention(q, k, rpb, 1, 1.0); b_.record()
        torch.cuda.synchronize()
        print("attention only:", min(a_.elapsed_time(b_) for a_, b_ in evs) * 1e3, "us")
elif what in ("conv8", "conv4", "conv1"):
    cin, cout = {"conv8": (8, 8), "conv4": (4, 8), "conv1": (1, 4)}[what]
    x = torch.randn(2, cin, *S, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.1
    b = torch.randn(cout, device=dev, generator=g) * 0.1
    st = torch.stack([x.double().sum((2, 3, 4)).flatten(), (x.double() ** 2).sum((2, 3, 4)).flatten()], 1).contiguous()
    fn = lambda: ops.conv3d(x, w, b, in_stats=st if cin > 1 else None, want_stats=True, act_out=cin == 1)
    nbytes = 2 * (cin + cout) * 4 * S[0] * S[1] * S[2]
elif what in ("conv128", "conv64_20", "conv32_40", "conv16_80"):
    cin, cout, shp = {"conv128": (128, 128, (10, 12, 10)), "conv64_20": (64, 64, (20, 24, 20)),
                      "conv32_40": (32, 32, (40, 48, 40)), "conv16_80": (16, 16, (80, 96, 80))}[what]
    x = torch.randn(2, cin, *shp, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.05
    b = torch.randn(cout, device=dev, generator=g) * 0.1
    st = torch.stack([x.double().sum((2, 3, 4)).flatten(), (x.double() ** 2).sum((2, 3, 4)).flatten()], 1).contiguous()
    fn = lambda: ops.conv3d(x, w, b, in_stats=st, want_stats=True)
    nbytes = 2 * (cin + cout) * 4 * shp[0] * shp[1] * shp[2]
    print(f"GFLOP {2 * 27 * cin * cout * 2 * shp[0] * shp[1] * shp[2] / 1e9:.2f}")
elif what == "warp8":
    src = torch.randn(1, 8, *S, device=dev, generator=g)
    flow = smooth_flow(S, 3.0)
    fn = lambda: ops.warp3d(src, flow)
    nbytes = (8 * 8 + 12) * S[0] * S[1] * S[2]
elif what == "proj":
    feat = torch.randn(2, 8, *S, device=dev, generator=g)
    w = torch.randn(6, 8, device=dev, generator=g)
    z = torch.zeros(6, device=dev)
    o = torch.ones(6, device=dev)
    fn = lambda: ops.proj_ln(feat, w, z, o, z)
    nbytes = 2 * (8 + 6) * 4 * S[0] * S[1] * S[2]
else:
    raise SystemExit(f"unknown kernel {what}")

with torch.no_grad():
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b_ in ev:
        flush.zero_()
        a.record()
        fn()
        b_.record()
    torch.cuda.synchronize()
ms = [a.elapsed_time(b_) for a, b_ in ev]
print(f"{what}: median {statistics.median(ms) * 1e3:.1f} us  min {min(ms) * 1e3:.1f} us  "
      f"algorithmic {nbytes / 1e6:.1f} MB -> {nbytes / statistics.median(ms) / 1e6:.0f} GB/s")
