"""Host-to-host registration loop: what `ModeT/infer.py:78-90` does per pair (upload the pair, run the
model, bring the warped image and the flow back), software-pipelined so the PCIe copies of pair i+1 /
i-1 overlap the kernels of pair i.

Three CUDA streams (upload, compute, download) and `depth` slots of device input / pinned host output
buffers; every hand-over is a CUDA event, the host only blocks when it hands a finished result to the
caller.  Pairs are independent, so this is also the unit that is replicated per GPU (one process per
GPU, `parallel.shard_range` picks each rank's pairs; no data-path collective)."""
from __future__ import annotations

from typing import Iterable, Iterator, List, Sequence, Tuple

import torch


class RegistrationPipeline:
    def __init__(self, model: torch.nn.Module, shape: Sequence[int], depth: int = 2, device=None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("RegistrationPipeline needs a CUDA model (smilecode_b200 has no CPU path)")
        self.depth = max(1, int(depth))
        D, H, W = (int(s) for s in shape)
        dev = self.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        mk = lambda c: torch.empty((1, c, D, H, W), dtype=torch.float32, device=dev)
        pin = lambda c: torch.empty((1, c, D, H, W), dtype=torch.float32).pin_memory()
        self.moving_d: List[torch.Tensor] = [mk(1) for _ in range(self.depth)]
        self.fixed_d: List[torch.Tensor] = [mk(1) for _ in range(self.depth)]
        self.moved_h: List[torch.Tensor] = [pin(1) for _ in range(self.depth)]
        self.flow_h: List[torch.Tensor] = [pin(3) for _ in range(self.depth)]
        ev = lambda: [torch.cuda.Event() for _ in range(self.depth)]
        self.in_ready, self.comp_done, self.out_done = ev(), ev(), ev()
        self.h2d_bytes = 2 * D * H * W * 4
        self.d2h_bytes = 4 * D * H * W * 4

    @torch.no_grad()
    def run(self, pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """pairs: iterable of (moving, fixed) CPU tensors [1,1,D,H,W] (pinned memory for true overlap).
        Yields (moved, flow) pinned CPU tensors in order; a yielded pair of buffers is reused `depth` pairs
        later, so consume (or copy) it before asking for that many more results."""
        compute = torch.cuda.current_stream(self.device)
        pending: List[int] = []
        keep = [None] * self.depth      # device outputs stay referenced until their download has been enqueued
        for i, (moving, fixed) in enumerate(pairs):
            s = i % self.depth
            if len(pending) == self.depth:              # slot s is about to be reused: hand its result out first
                j = pending.pop(0)
                self.out_done[j].synchronize()
                yield self.moved_h[j], self.flow_h[j]
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.comp_done[s])   # the kernels that last read this slot's inputs are done
                self.moving_d[s].copy_(moving, non_blocking=True)
                self.fixed_d[s].copy_(fixed, non_blocking=True)
                self.in_ready[s].record(self.s_in)
            compute.wait_event(self.in_ready[s])
            moved, flow = self.model(self.moving_d[s], self.fixed_d[s])
            self.comp_done[s].record(compute)
            keep[s] = (moved, flow)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.comp_done[s])
                self.moved_h[s].copy_(moved, non_blocking=True)
                self.flow_h[s].copy_(flow, non_blocking=True)
                moved.record_stream(self.s_out)
                flow.record_stream(self.s_out)
                self.out_done[s].record(self.s_out)
            pending.append(s)
        for j in pending:
            self.out_done[j].synchronize()
            yield self.moved_h[j], self.flow_h[j]
