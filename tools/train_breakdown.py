#!/usr/bin/env python
"""Per-kernel CUDA-event breakdown of one fp32 training step at 160x192x160 (same setup as bench.py's train leg)."""
import sys

import torch

sys.path.insert(0, ".")
from smilecode_b200 import _lib, models  # noqa: E402
from smilecode_b200.synth import make_pair, randomize_weights  # noqa: E402
from smilecode_b200.train import Trainer  # noqa: E402

dev = torch.device("cuda")
SHAPE = (160, 192, 160)
model = models.ModeT(SHAPE, head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1)
randomize_weights(model, seed=1234)
model = model.to(dev)
import os
model.conv_precision = os.environ.get("SMILE_TRAIN_DTYPE", "fp32")
tr = Trainer(model, lr=1e-4)
moving, fixed = [t.to(dev) for t in make_pair(SHAPE, batch=1, seed=24)]
for _ in range(2):
    tr.step(moving, fixed)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); tr.step(moving, fixed); b.record(); torch.cuda.synchronize()
print(f"step {a.elapsed_time(b):.2f} ms")
_lib.profile_start()
tr.step(moving, fixed)
prof = _lib.profile_stop()
rows = sorted(((v[1], v[0], k) for k, v in prof.items()), reverse=True)
tot = sum(r[0] for r in rows)
agg = {}
for ms, calls, name in rows:
    agg[name.split("[")[0]] = agg.get(name.split("[")[0], 0.0) + ms
print(f"sum of C-ABI kernels {tot:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
    print(f"  {v:8.3f} ms {100 * v / tot:5.1f}%  {k}")
print("top launches:")
for ms, calls, name in rows[:22]:
    print(f"  {ms:8.3f} ms x{calls:<2d} {name}")
