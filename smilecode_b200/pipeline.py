"""Host-to-host registration loop: what `ModeT/infer.py:78-92` does per pair (upload the pair, run the model, bring the
result back), software-pipelined so the PCIe copies of pair i+1 / i-1 overlap the kernels of pair i.

Three CUDA streams (upload, compute, download), `depth` device input slots and `2 * depth` pinned host output slots;
every hand-over is a CUDA event, the host only blocks when it hands a finished result to the caller.  Pairs are
independent, so this is also the unit that is replicated per GPU (one process per GPU, `parallel.shard_range` picks each
rank's pairs; no data-path collective).

What comes back is selectable, because on a multi-GPU box the host side of the PCIe copies is the shared resource
(8 ranks x 118 MB per pair saturate the host at ~118 GB/s):
  outputs=("flow",)            the reference loop: infer.py:89 takes only `flow` to the host (default; 59 MB per pair)
  outputs=("moved", "flow")    both results of ModeT.forward (79 MB per pair)
  outputs=()  + `reduce=`      nothing but what `reduce(moved, flow, extra)` returns, evaluated on the device: e.g. the Dice /
                               Jacobian numbers of infer.py:87-92 through smilecode_b200.metrics (a few bytes per pair)
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


class RegistrationPipeline:
    def __init__(self, model: torch.nn.Module, shape: Sequence[int], depth: int = 3, device=None,
                 outputs: Sequence[str] = ("flow",), reduce: Optional[Callable] = None, cuda_graph: bool = True):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("RegistrationPipeline needs a CUDA model (smilecode_b200 has no CPU path)")
        for o in outputs:
            if o not in ("moved", "flow"):
                raise ValueError(f"unknown output {o!r} (choose from 'moved', 'flow')")
        self.depth = max(1, int(depth))
        self.outputs = tuple(outputs)
        self.reduce = reduce
        # one CUDA graph of the forward per device slot (graph.GraphedForward), captured at the slot's first use; the
        # model's weights must stay as they are while the pipeline lives (cuda_graph=False launches kernel by kernel)
        self.cuda_graph = bool(cuda_graph)
        self._graphs: List = [None] * self.depth
        D, H, W = (int(s) for s in shape)
        dev = self.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        mk = lambda c: torch.empty((1, c, D, H, W), dtype=torch.float32, device=dev)
        pin = lambda c: torch.empty((1, c, D, H, W), dtype=torch.float32).pin_memory()
        self.moving_d: List[torch.Tensor] = [mk(1) for _ in range(self.depth)]
        self.fixed_d: List[torch.Tensor] = [mk(1) for _ in range(self.depth)]
        # 2 * depth host slots: pair j is handed out while pair j + depth is being enqueued, and its slot is written
        # again by pair j + 2 * depth, i.e. after `depth` further results have been requested (with `depth`
        # slots it was overwritten after ONE more request)
        self.nhost = 2 * self.depth
        chans = {"moved": 1, "flow": 3}
        self.out_h = [{o: pin(chans[o]) for o in self.outputs} for _ in range(self.nhost)]
        self.red_h: List[Optional[torch.Tensor]] = [None] * self.nhost     # pinned landing buffers of `reduce` results
        ev = lambda n: [torch.cuda.Event() for _ in range(n)]
        self.in_ready, self.comp_done = ev(self.depth), ev(self.depth)
        self.out_done = ev(self.nhost)
        self.h2d_bytes = 2 * D * H * W * 4
        self.d2h_bytes = sum(chans[o] for o in self.outputs) * D * H * W * 4

    @torch.no_grad()
    def run(self, pairs: Iterable) -> Iterator:
        """pairs: iterable of (moving, fixed) or (moving, fixed, extra) with moving / fixed CPU tensors [1,1,D,H,W]
        (pinned memory for true overlap); `extra` is passed through to `reduce` untouched.
        Yields, in order, one item per pair: the requested pinned host tensors as a tuple in the order of `outputs`
        (a single tensor if only one was requested), followed by the value of `reduce` when one was given.
        A yielded host buffer stays valid until `depth` further items have been requested."""
        compute = torch.cuda.current_stream(self.device)
        pending: List[Tuple[int, object]] = []          # (host slot, reduce result) in flight
        for i, item in enumerate(pairs):
            moving, fixed = item[0], item[1]
            extra = item[2] if len(item) > 2 else None
            s, hs = i % self.depth, i % self.nhost
            if len(pending) == self.depth:              # keeps at most `depth` pairs in flight
                yield self._finish(*pending.pop(0))
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.comp_done[s])   # the kernels that last read this slot's inputs are done
                self.moving_d[s].copy_(moving, non_blocking=True)
                self.fixed_d[s].copy_(fixed, non_blocking=True)
                self.in_ready[s].record(self.s_in)
            compute.wait_event(self.in_ready[s])
            if self.cuda_graph:
                if self._graphs[s] is None:
                    from .graph import GraphedForward
                    self._graphs[s] = GraphedForward(self.model, self.moving_d[s], self.fixed_d[s])
                    compute.wait_event(self.in_ready[s])
                if i >= self.depth:     # the graph's output tensors are reused: their last download must be complete
                    compute.wait_event(self.out_done[(i - self.depth) % self.nhost])
                moved, flow = self._graphs[s].replay()
            else:
                moved, flow = self.model(self.moving_d[s], self.fixed_d[s])
            red = self.reduce(moved, flow, extra) if self.reduce is not None else None
            self.comp_done[s].record(compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.comp_done[s])
                for name, src in (("moved", moved), ("flow", flow)):
                    if name in self.out_h[hs]:
                        self.out_h[hs][name].copy_(src, non_blocking=True)
                        if not self.cuda_graph:
                            src.record_stream(self.s_out)
                if isinstance(red, torch.Tensor):      # device result of `reduce`: same asynchronous route to the host
                    if self.red_h[hs] is None or self.red_h[hs].shape != red.shape or self.red_h[hs].dtype != red.dtype:
                        self.red_h[hs] = torch.empty(red.shape, dtype=red.dtype).pin_memory()
                    self.red_h[hs].copy_(red, non_blocking=True)
                    red.record_stream(self.s_out)
                    red = self.red_h[hs]
                self.out_done[hs].record(self.s_out)
            pending.append((hs, red))
        while pending:
            yield self._finish(*pending.pop(0))

    def _finish(self, hs: int, red):
        self.out_done[hs].synchronize()
        outs = tuple(self.out_h[hs][o] for o in self.outputs)
        if self.reduce is not None:
            return (*outs, red) if outs else red
        return outs[0] if len(outs) == 1 else outs
