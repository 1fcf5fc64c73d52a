"""CUDA-graph replay of the inference forward for fixed input buffers.

`ModeT.forward` at 160x192x160 is ~50 dependent kernel launches through the C ABI; on the coarse pyramid levels the kernels
run for 10-50 microseconds each, so launch latency and the Python / ctypes dispatch between them is visible (four pairs per
launch are 13 % faster per pair than one).  Capturing the forward once per input-buffer pair and replaying it removes that
host work from the steady state.  Nothing is traced or compiled: the graph is the same sequence of hand-written kernels, with
their arguments (device pointers, TMA descriptors) frozen -- which is why the inputs must live at fixed addresses
(`pipeline.RegistrationPipeline` replays one graph per device slot).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


class GraphedForward:
    """graph = GraphedForward(model, moving_buf, fixed_buf);  moved, flow = graph.replay()

    `moving_buf` / `fixed_buf` are device tensors whose CONTENTS the caller updates between replays; the returned tensors
    are owned by the graph and overwritten by the next replay.  The model must be in inference mode (no autograd) and its
    weights must not be replaced while the graph is alive (prepared tensor-core weight blocks are baked in)."""

    def __init__(self, model: torch.nn.Module, moving: torch.Tensor, fixed: torch.Tensor, warmup: int = 2):
        if not (moving.is_cuda and fixed.is_cuda):
            raise _lib.SmileError("GraphedForward needs CUDA input buffers")
        self.model, self.moving, self.fixed = model, moving, fixed
        cur = torch.cuda.current_stream(moving.device)
        side = torch.cuda.Stream(moving.device)
        side.wait_stream(cur)
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # prepares cached weights / function attributes outside the capture
                model(moving, fixed)
        cur.wait_stream(side)
        torch.cuda.synchronize(moving.device)
        l0 = _lib.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.moved, self.flow = model(moving, fixed)
        self.launches = _lib.LAUNCHES - l0           # kernels inside one replay (bench.py's gpu_launches counts them)
        _lib.LAUNCHES = l0                           # capture enqueued nothing

    def replay(self) -> Tuple[torch.Tensor, torch.Tensor]:
        self.graph.replay()
        _lib.LAUNCHES += self.launches
        return self.moved, self.flow
