#!/usr/bin/env python
"""Host<->device copy bandwidth per rank, alone and with all ranks copying at once (where does the
end-to-end path saturate at 8 GPUs?).

    python tools/pcie_probe.py                      # one GPU
    torchrun --nproc-per-node 8 tools/pcie_probe.py # all ranks at once; rank 0 prints the table

Per rank: 256 MiB pinned buffers, 10 repetitions each of H2D, D2H and both directions concurrently on two streams; CUDA
events; the all-ranks numbers are taken between barriers so that the copies of all ranks overlap."""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = 256 << 20
h_in = torch.empty(NB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(NB, dtype=torch.uint8).pin_memory()
d_a = torch.empty(NB, dtype=torch.uint8, device=dev)
d_b = torch.empty(NB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 10


def run(mode):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a)
    s2.wait_event(a)
    for _ in range(REPS):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    nbytes = REPS * NB * (2 if mode == "both" else 1)
    return nbytes / ms / 1e6      # GB/s


for _ in range(2):
    run("both")
res = {m: run(m) for m in ("h2d", "d2h", "both")}
if world > 1:
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], device=dev, dtype=torch.float64)
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    if rank == 0:
        rows = [[float(x) for x in v] for v in allv]
        out = {"world": world, "per_rank_GBs_h2d_d2h_both": rows,
               "aggregate_GBs": {"h2d": sum(r[0] for r in rows), "d2h": sum(r[1] for r in rows), "both": sum(r[2] for r in rows)},
               "min_rank_GBs": {"h2d": min(r[0] for r in rows), "d2h": min(r[1] for r in rows), "both": min(r[2] for r in rows)}}
        print("PCIE " + json.dumps(out))
    dist.destroy_process_group()
else:
    print("PCIE " + json.dumps({"world": 1, "GBs": res}))
