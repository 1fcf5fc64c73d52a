#!/usr/bin/env python
"""CUDA-event times of the infer.py evaluation kernels at 160x192x160 next to the reference's host path
(device->host copy of the flow + numpy jacobian + numpy dice)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import metrics_oracle as morc   # host baseline / checker  # noqa: E402
from smilecode_b200 import metrics  # noqa: E402

S = (160, 192, 160)
g = torch.Generator(device="cuda").manual_seed(3)
c = torch.randn(1, 3, 10, 12, 10, device="cuda", generator=g) * 3
flow = torch.nn.functional.interpolate(c, size=S, mode="trilinear", align_corners=True).contiguous()
seg_m = torch.randint(0, 55, (1, 1, *S), device="cuda", generator=g).float()
seg_f = torch.randint(0, 55, (1, 1, *S), device="cuda", generator=g).float()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2] * 1e3


print(f"nearest warp      {timed(lambda: metrics.warp3d_nearest(seg_m, flow)):8.1f} us")
print(f"dice counts       {timed(lambda: metrics.dice_counts(seg_m, seg_f)):8.1f} us")
print(f"jacobian (count)  {timed(lambda: metrics.jacobian_determinant_vxm(flow, want_det=False)):8.1f} us")
print(f"jacobian (+det)   {timed(lambda: metrics.jacobian_determinant_vxm(flow, want_det=True)):8.1f} us")
t0 = time.perf_counter()
fh = flow.cpu().numpy()[0]
t1 = time.perf_counter()
det = morc.jacobian_determinant_vxm(fh)
t2 = time.perf_counter()
d = morc.dice_val_VOI(seg_m.long().cpu().numpy()[0, 0], seg_f.long().cpu().numpy()[0, 0])
t3 = time.perf_counter()
print(f"reference host path: flow D2H {1e3 * (t1 - t0):.1f} ms, numpy jacobian {1e3 * (t2 - t1):.0f} ms, numpy dice (+ copies) {1e3 * (t3 - t2):.0f} ms")
_, nonpos = metrics.jacobian_determinant_vxm(flow, want_det=False)
assert int(nonpos.item()) == int(np.sum(det <= 0)) and metrics.dice_val_VOI(seg_m, seg_f) == d
print("results identical to the host path")
