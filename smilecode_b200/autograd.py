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


class ConvINLReLU(Function):
    """ConvInsBlock (models.py:135-151): act = LeakyReLU(InstanceNorm(Conv3d(x))), optionally also AvgPool3d(2)(act)
    (models.py:198).  Saves x, weight, act and the fp64 statistics; the raw conv output is not kept."""

    @staticmethod
    def forward(ctx, x, weight, bias, pool):
        x, weight, bias = _c(x), _c(weight), _c(bias)
        raw, stats = ops.conv3d(x, weight, bias, want_stats=True)
        act, pooled = ops.instnorm_lrelu_pool(raw, stats, pool=bool(pool), inplace=True)
        ctx.save_for_backward(x, weight, act, stats)
        ctx.pool = bool(pool)
        ctx.prec = ops.current_conv_precision()      # the data-gradient pass runs at the forward's precision
        if pool:
            return act, pooled
        return act

    @staticmethod
    def backward(ctx, g_act, g_pooled=None):
        x, weight, act, stats = ctx.saved_tensors
        g = _c(g_act)
        if ctx.pool and g_pooled is not None:
            g = g.clone() if g is not None else torch.zeros_like(act)
            ops.avgpool2_bwd_add(_c(g_pooled), g)
        d_raw = ops.in_lrelu_bwd(g, act, stats, mode=0)
        with ops.conv_precision(ctx.prec):
            dx, dw, db = ops.conv3d_bwd(d_raw, x, weight, need_x=ctx.needs_input_grad[0])
        return dx, dw, db, None


class ConvLReLU(Function):
    """ConvBlock (models.py:119-133): LeakyReLU(Conv3d(x))."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight, bias = _c(x), _c(weight), _c(bias)
        act, _ = ops.conv3d(x, weight, bias, act_out=True)
        ctx.save_for_backward(x, weight, act)
        ctx.prec = ops.current_conv_precision()
        return act

    @staticmethod
    def backward(ctx, g):
        x, weight, act = ctx.saved_tensors
        d_raw = ops.in_lrelu_bwd(_c(g), act, None, mode=1)
        with ops.conv_precision(ctx.prec):
            dx, dw, db = ops.conv3d_bwd(d_raw, x, weight, need_x=ctx.needs_input_grad[0])
        return dx, dw, db


class Conv(Function):
    """Plain Conv3d(k=3, s=1, p=1) (CWM's last conv, models.py:253)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight, bias = _c(x), _c(weight), _c(bias)
        ctx.save_for_backward(x, weight)
        ctx.prec = ops.current_conv_precision()
        return ops.conv3d(x, weight, bias)[0]

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        with ops.conv_precision(ctx.prec):
            dx, dw, db = ops.conv3d_bwd(_c(g), x, weight, need_x=ctx.needs_input_grad[0])
        return dx, dw, db


class NCCLoss(Function):
    """NCC_vxm.forward(y_true, y_pred) (losses.py:43-95); gradient w.r.t. y_true (the warped image, train.py:126)."""

    @staticmethod
    def forward(ctx, y_true, y_pred, win):
        y_true, y_pred = _c(y_true), _c(y_pred)
        ctx.save_for_backward(y_true, y_pred)
        ctx.win = int(win)
        return ops.ncc_vxm(y_true, y_pred, ctx.win)

    @staticmethod
    def backward(ctx, g):
        y_true, y_pred = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("NCCLoss: gradient w.r.t. the second argument is not implemented")
        gs = g.detach().to(torch.float32).reshape(1).contiguous()
        return ops.ncc_vxm_bwd(y_true, y_pred, gs, ctx.win), None, None


class Grad3dLoss(Function):
    """Grad3d(penalty='l2') (losses.py:16-31)."""

    @staticmethod
    def forward(ctx, flow):
        flow = _c(flow)
        ctx.save_for_backward(flow)
        return ops.grad3d_l2(flow)

    @staticmethod
    def backward(ctx, g):
        (flow,) = ctx.saved_tensors
        gs = g.detach().to(torch.float32).reshape(1).contiguous()
        return ops.grad3d_l2_bwd(flow, gs)
