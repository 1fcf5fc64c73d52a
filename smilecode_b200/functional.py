"""Drop-in for the reference's `ModeT-cu/functional.py`: `modetqkrpb_cu(query, key, rpb)` with the same
autograd contract (forward saves q, k; backward returns d_query, d_key_padded, d_rpb), running the
sm_100a kernels behind `smile_modet_qkrpb_{fwd,bwd}` instead of the `modet` pybind extension
(ModeT-cu/functional.py:5-28, ModeT-cu/modet/modet.cpp:34-37)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops


class ModeTFunction(Function):
    @staticmethod
    def forward(ctx, query, key, rpb):
        query = query.detach().contiguous()
        key = key.detach().contiguous()
        bias = rpb is not None
        with torch.no_grad():
            attn = ops.modet_qkrpb_fwd(query, key, rpb.detach().contiguous() if bias else None)
        ctx.save_for_backward(query, key)
        ctx.bias = bias
        return attn

    @staticmethod
    def backward(ctx, grad_out):
        query, key = ctx.saved_tensors
        with torch.no_grad():
            dq, dk, drpb = ops.modet_qkrpb_bwd(grad_out.detach().contiguous(), query, key, ctx.bias)
        return dq, dk, drpb


def modetqkrpb_cu(query, key, rpb):
    return ModeTFunction.apply(query, key, rpb)
