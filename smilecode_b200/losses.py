"""Same class names / call signatures as the reference's `ModeT/losses.py` (`NCC_vxm`, `Grad3d`), forward
and backward on the sm_100a kernels (`smile_ncc_vxm_{fwd,bwd}`, `smile_grad3d_l2_{fwd,bwd}`).  NCC_vxm is
differentiated w.r.t. its first argument (the warped image, as ModeT/train.py:126 calls it)."""
from __future__ import annotations

import torch

from . import ops


class Grad3d(torch.nn.Module):
    """Reference ModeT/losses.py:6-31.  Only penalty='l2' (what train.py:47 uses) is on the hot path."""

    def __init__(self, penalty="l1", loss_mult=None):
        super().__init__()
        self.penalty = penalty
        self.loss_mult = loss_mult

    def forward(self, y_pred, y_true=None):
        if self.penalty != "l2":
            raise NotImplementedError("Grad3d: only penalty='l2' is implemented (ModeT/train.py:47)")
        if y_pred.requires_grad and torch.is_grad_enabled():
            from .autograd import Grad3dLoss
            grad = Grad3dLoss.apply(y_pred)
        else:
            grad = ops.grad3d_l2(y_pred)
        return grad * self.loss_mult if self.loss_mult is not None else grad


class NCC_vxm(torch.nn.Module):
    """Reference ModeT/losses.py:34-95 (3-D volumes, cubic window, default 9)."""

    def __init__(self, win=None):
        super().__init__()
        self.win = win

    def forward(self, y_true, y_pred):
        win = 9 if self.win is None else (self.win[0] if isinstance(self.win, (list, tuple)) else int(self.win))
        if torch.is_grad_enabled() and (y_true.requires_grad or y_pred.requires_grad):
            from .autograd import NCCLoss
            return NCCLoss.apply(y_true, y_pred, win)
        return ops.ncc_vxm(y_true, y_pred, win)
