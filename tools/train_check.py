#!/usr/bin/env python
"""Full-size training step timing (and, under torchrun, the data-parallel all-reduce check).
    python tools/train_check.py [steps]            # 1 GPU
    torchrun --nproc-per-node 2 tools/train_check.py [steps]
"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from smilecode_b200 import _lib, models
from smilecode_b200.synth import make_pair, randomize_weights
from smilecode_b200.train import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
shape = tuple(int(x) for x in os.environ.get("SHAPE", "160x192x160").split("x"))
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = models.ModeT(shape, head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1)
randomize_weights(model, seed=1234)
model = model.cuda()
tr = Trainer(model, lr=1e-4, distributed=world > 1)
moving, fixed = (t.cuda() for t in make_pair(shape, batch=1, seed=24 + rank))
for _ in range(2):
    loss, ncc, reg = tr.step(moving, fixed)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
l0 = _lib.LAUNCHES
t0 = time.perf_counter()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    loss, ncc, reg = tr.step(moving, fixed)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
if world > 1:
    # replicas must hold identical parameters after identical averaged updates
    flat = tr.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    diff = float((flat - ref).abs().max())
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
    if rank == 0:
        print(f"replica parameter divergence after {steps + 2} steps: {diff:.3e}")
if rank == 0:
    print(f"train step {shape} fp32 B=1/GPU x{world}: {ms:.2f} ms/step -> {world * 1e3 / ms:.2f} pairs/s; "
          f"loss {float(loss):.5f} ncc {float(ncc):.5f} reg {float(reg):.5f}; launches/step {(_lib.LAUNCHES - l0) // steps}; "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if os.environ.get("PROFILE") and rank == 0:
    _lib.profile_start()
    tr.step(moving, fixed)
    prof = _lib.profile_stop()
    rows = sorted(((v[1], v[0], k) for k, v in prof.items()), reverse=True)
    tot = sum(r[0] for r in rows)
    for ms_, calls, name in rows[:40]:
        print(f"  {ms_:9.3f} ms  {100 * ms_ / tot:5.1f}%  x{calls:<3d} {name}")
    print(f"  total of kernels {tot:.3f} ms")
if world > 1:
    dist.destroy_process_group()
