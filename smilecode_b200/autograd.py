"""torch.autograd glue for the training path: every Function's forward AND backward run the hand-written
sm_100a kernels through the C ABI (`ops.*` / `ops.*_bwd`); autograd only chains them (the role
`ModeT-cu/functional.py:5-28` plays for the reference's single custom op)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops


def _c(t):
    return None if t is None else t.detach().contiguous()


class Warp(Function):
    """SpatialTransformer (ModeT/models.py:49-67)."""

    @staticmethod
    def forward(ctx, src, flow):
        src, flow = _c(src), _c(flow)
        ctx.save_for_backward(src, flow)
        return ops.warp3d(src, flow)

    @staticmethod
    def backward(ctx, g):
        src, flow = ctx.saved_tensors
        d_src, d_flow = ops.warp3d_bwd(_c(g), src, flow, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return d_src, d_flow


class Upsample2x(Function):
    """premul * nn.Upsample(scale_factor=2, trilinear, align_corners=True) (models.py:354)."""

    @staticmethod
    def forward(ctx, x, premul):
        ctx.premul = float(premul)
        return ops.upsample2x(_c(x), ctx.premul)

    @staticmethod
    def backward(ctx, g):
        return ops.upsample2x_bwd(_c(g), ctx.premul), None


class Attention(Function):
    """ModeTransformer.forward (models.py:308-334)."""

    @staticmethod
    def forward(ctx, q, k, rpb, heads, scale):
        q, k, rpb = _c(q), _c(k), _c(rpb)
        ctx.save_for_backward(q, k, rpb)
        ctx.heads, ctx.scale = int(heads), float(scale)
        return ops.modet_attention(q, k, rpb, ctx.heads, ctx.scale)

    @staticmethod
    def backward(ctx, g):
        q, k, rpb = ctx.saved_tensors
        dq, dk, drpb = ops.modet_attention_bwd(_c(g), q, k, rpb, ctx.heads, ctx.scale)
        return dq, dk, drpb, None, None


class ProjLN(Function):
    """ProjectionLayer.forward (models.py:238-241)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, gamma, beta, eps):
        feat, weight, bias, gamma, beta = _c(feat), _c(weight), _c(bias), _c(gamma), _c(beta)
        ctx.save_for_backward(feat, weight, bias, gamma)
        ctx.eps = float(eps)
        return ops.proj_ln(feat, weight, bias, gamma, beta, ctx.eps)

    @staticmethod
    def backward(ctx, g):
        feat, weight, bias, gamma = ctx.saved_tensors
        dfeat, dw, db, dg, dbeta = ops.proj_ln_bwd(_c(g), feat, weight, bias, gamma, ctx.eps, ctx.needs_input_grad[0])
        return dfeat, dw, db, dg, dbeta, None


class CwmFuse(Function):
    """CWM tail (models.py:268-275)."""

    @staticmethod
    def forward(ctx, fields, logits):
        fields, logits = _c(fields), _c(logits)
        ctx.save_for_backward(fields, logits)
        return ops.cwm_fuse(fields, logits)

    @staticmethod
    def backward(ctx, g):
        fields, logits = ctx.saved_tensors
        return ops.cwm_fuse_bwd(_c(g), fields, logits)
