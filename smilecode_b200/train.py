"""Training step of the reference loop (`ModeT/train.py:114-133`) on the sm_100a kernels:
forward -> NCC_vxm + Grad3d('l2') -> backward (hand-written kernels chained by autograd) -> optional
data-parallel gradient all-reduce (one flat bucket, `parallel.FlatGradAllReduce`) -> fused Adam(amsgrad) update
of the flat parameter buffer (`smile_adam_amsgrad_step`), with the reference's polynomial learning-rate decay
(train.py:166-168)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import losses, ops
from .parallel import FlatGradAllReduce, flat_layout


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weights: Sequence[float] = (1.0, 1.0),
                 betas=(0.9, 0.999), eps: float = 1e-8, distributed: bool = False):
        self.model = model
        self.lr0 = float(lr)
        self.weights = tuple(float(w) for w in weights)
        self.betas, self.eps = betas, float(eps)
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        # flat parameter / gradient / optimizer-state buffers: parameters become views into `flat`
        # (16-byte aligned starts, zero padding: parallel.flat_layout -- the gradient bucket uses the same offsets)
        offs, self.numel = flat_layout(self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        for p, off in zip(self.params, offs):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.sync = FlatGradAllReduce(self.params)          # .grad tensors become views into sync.bucket
        self.distributed = bool(distributed)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.max_exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        self.ncc = losses.NCC_vxm()
        self.grad = losses.Grad3d(penalty="l2")

    def lr_at(self, epoch: int, max_epoch: int, power: float = 0.9) -> float:
        """adjust_learning_rate of ModeT/train.py:166-168."""
        return round(self.lr0 * (1 - epoch / max_epoch) ** power, 8)

    def step(self, moving: torch.Tensor, fixed: torch.Tensor, lr: Optional[float] = None):
        """One iteration of train.py:118-133.  Returns (loss, ncc, reg) as device scalars (no host sync here;
        the reference's `loss.item()` stays with the caller)."""
        self.model.train()
        self.sync.zero_()
        moved, flow = self.model(moving, fixed)
        l_ncc = self.ncc(moved, fixed) * self.weights[0]
        l_reg = self.grad(flow, fixed) * self.weights[1]
        loss = l_ncc + l_reg
        loss.backward()
        if self.distributed:
            self.sync.allreduce_mean_()
        self.step_count += 1
        ops.adam_amsgrad_step(self.flat, self.sync.bucket, self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq,
                              self.lr0 if lr is None else lr, self.betas[0], self.betas[1], self.eps, self.step_count)
        return loss.detach(), l_ncc.detach(), l_reg.detach()
