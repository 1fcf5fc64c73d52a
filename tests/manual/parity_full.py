#!/usr/bin/env python
"""Full-size (160x192x160) end-to-end parity report: our CUDA forward vs the CPU oracle in fp32 and fp64.
Prints |ours - ref32|, |ours - ref64| and the reference's own fp32 noise floor |ref32 - ref64|."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import modet_oracle as orc          # checker
from smilecode_b200 import models
from smilecode_b200.synth import make_pair

shape = tuple(int(x) for x in sys.argv[1].split("x")) if len(sys.argv) > 1 else (160, 192, 160)
heads = [8, 4, 2, 1, 1]
torch.set_num_threads(16)
sd = orc.synth_state_dict(seed=1234, num_heads=heads)
moving, fixed = make_pair(shape, batch=1, seed=24)
model = models.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
model.load_state_dict(sd, strict=False)
model = model.cuda().eval()
with torch.no_grad():
    moved, flow = model(moving.cuda(), fixed.cuda())
    moved, flow = moved.cpu(), flow.cpu()
    t0 = time.time()
    m32, f32 = orc.modet_forward(moving, fixed, sd, num_heads=heads, scale=1.0, library_ops=True)
    t1 = time.time()
    sd64 = {k: v.double() for k, v in sd.items()}
    m64, f64 = orc.modet_forward(moving.double(), fixed.double(), sd64, num_heads=heads, scale=1.0, library_ops=True)
    t2 = time.time()
mx = lambda a, b: float((a.double() - b.double()).abs().max())
print(f"shape {shape}: |flow| max {float(f64.abs().max()):.3f}  (cpu fp32 {t1 - t0:.1f} s, fp64 {t2 - t1:.1f} s)")
print(f"flow : |ours-ref32| {mx(flow, f32):.3e}  |ours-ref64| {mx(flow, f64):.3e}  |ref32-ref64| {mx(f32, f64):.3e}")
print(f"moved: |ours-ref32| {mx(moved, m32):.3e}  |ours-ref64| {mx(moved, m64):.3e}  |ref32-ref64| {mx(m32, m64):.3e}")
d = (flow.double() - f64).abs()
print(f"flow error vs ref64: mean {float(d.mean()):.3e}  p99.9 {float(d.flatten()[::7].kthvalue(int(0.999 * d.flatten()[::7].numel())).values):.3e}")
