"""Tensor-level operators of the ModeT hot path: thin wrappers that validate torch tensors,
allocate outputs through torch's caching allocator and enqueue the sm_100a kernels on torch's
current stream through the C ABI (`_lib.call`).  PyTorch is plumbing here (memory + streams);
every arithmetic step runs in libsmilecode_b200.so.  No fallbacks.

Reference functions replaced are cited per operator (paths under the reference tree).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import SmileError, call

Tensor = torch.Tensor
IN_EPS = 1e-5  # nn.InstanceNorm3d / nn.LayerNorm default eps (ModeT/models.py:144, 233)


def _chk(t: Tensor, name: str, ndim: int, dtype=torch.float32) -> Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise SmileError(f"{name} must be a CUDA tensor (smilecode_b200 has no CPU path)")
    if t.dtype != dtype:
        raise SmileError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() != ndim:
        raise SmileError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    if t.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError(
            f"{name} requires grad: smilecode_b200.ops are the raw forward kernels; differentiable calls go through "
            "smilecode_b200.autograd (ModeT.forward does this by itself when grad is enabled), or wrap the call in "
            "torch.no_grad()")
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


class stats_arena:
    """One zero-filled fp64 buffer per forward pass that the InstanceNorm statistics of every conv3d are carved from
    (one fill kernel instead of one per layer: sixteen 2-microsecond launches per ModeT forward).

        with ops.stats_arena(device):
            ... ops.conv3d(..., want_stats=True) ...
    """
    _active = None

    def __init__(self, device, rows: int = 4096):
        self.device, self.rows = device, rows

    def __enter__(self):
        self.buf = torch.zeros((self.rows, 2), device=self.device, dtype=torch.float64)
        self.used = 0
        self._prev, stats_arena._active = stats_arena._active, self
        return self

    def __exit__(self, *exc):
        stats_arena._active = self._prev
        return False

    @staticmethod
    def take(rows: int, device) -> Tensor:
        a = stats_arena._active
        if a is None or a.buf.device != device or a.used + rows > a.rows:
            return torch.zeros((rows, 2), device=device, dtype=torch.float64)
        out = a.buf[a.used:a.used + rows]
        a.used += rows
        return out


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def modet_attention(q: Tensor, k: Tensor, rpb: Optional[Tensor], heads: int, scale: float) -> Tensor:
    """ModeTransformer.forward (ModeT/models.py:308-334).  q,k [B,D,H,W,C] -> [B,3*heads,D,H,W]."""
    q = _chk(q, "q", 5)
    k = _chk(k, "k", 5)
    if q.shape != k.shape:
        raise SmileError(f"q {tuple(q.shape)} and k {tuple(k.shape)} differ")
    B, D, H, W, C = q.shape
    if C % heads:
        raise SmileError(f"channels {C} not divisible by heads {heads}")
    if rpb is not None:
        rpb = _chk(rpb, "rpb", 4)
        if tuple(rpb.shape) != (heads, 3, 3, 3):
            raise SmileError(f"rpb must be [{heads},3,3,3], got {tuple(rpb.shape)}")
    out = torch.empty((B, 3 * heads, D, H, W), device=q.device, dtype=torch.float32)
    call("smile_modet_attn_fwd", q.data_ptr(), k.data_ptr(), _ptr(rpb), out.data_ptr(), B, D, H, W, heads, C // heads,
         float(scale), _stream(), label=f"[h{heads} {D}x{H}x{W}]")
    return out


def warp3d(src: Tensor, flow: Tensor) -> Tensor:
    """SpatialTransformer.forward, bilinear mode (ModeT/models.py:49-67)."""
    src = _chk(src, "src", 5)
    flow = _chk(flow, "flow", 5)
    B, C, D, H, W = src.shape
    if tuple(flow.shape) != (B, 3, D, H, W):
        raise SmileError(f"flow must be [{B},3,{D},{H},{W}], got {tuple(flow.shape)}")
    out = torch.empty_like(src)
    call("smile_warp3d_fwd", src.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, D, H, W, _stream(),
         label=f"[c{C} {D}x{H}x{W}]")
    return out


def upsample2x(x: Tensor, premul: float = 1.0) -> Tensor:
    """premul * nn.Upsample(scale_factor=2, trilinear, align_corners=True)(x) (ModeT/models.py:354)."""
    x = _chk(x, "x", 5)
    B, C, D, H, W = x.shape
    out = torch.empty((B, C, 2 * D, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
    call("smile_upsample2x_fwd", x.data_ptr(), out.data_ptr(), B, C, D, H, W, float(premul), _stream(),
         label=f"[c{C} {D}x{H}x{W}]")
    return out


def flow_compose(flow: Tensor, w: Tensor, postmul: float = 1.0) -> Tensor:
    """postmul * (SpatialTransformer(flow, w) + w) (ModeT/models.py:392, 398, 403, 408)."""
    flow = _chk(flow, "flow", 5)
    w = _chk(w, "w", 5)
    B, three, D, H, W = flow.shape
    if three != 3 or w.shape != flow.shape:
        raise SmileError(f"flow {tuple(flow.shape)} / w {tuple(w.shape)} must both be [B,3,D,H,W]")
    out = torch.empty_like(flow)
    call("smile_flow_compose_fwd", flow.data_ptr(), w.data_ptr(), out.data_ptr(), B, D, H, W, float(postmul), _stream(),
         label=f"[{D}x{H}x{W}]")
    return out


def modet_fused(q: Tensor, k: Tensor, rpb: Optional[Tensor], flow_in: Tensor, moving: Optional[Tensor], scale: float,
                postmul: float = 1.0, ln_gamma: Optional[Tensor] = None,
                ln_beta: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """heads==1 level in one kernel (ModeT/models.py:401-403, 406-410):
    w = mdt(q,k); flow_out = postmul*(T(flow_in,w)+w); moved = T(moving, flow_out) if moving is given.
    ln_gamma / ln_beta: the LayerNorm affine parameters of the ProjectionLayer that produced q AND k (both or neither);
    they let the kernel bound the logits and skip the running maximum of the softmax (see include/smilecode_b200.h)."""
    q = _chk(q, "q", 5)
    k = _chk(k, "k", 5)
    flow_in = _chk(flow_in, "flow_in", 5)
    B, D, H, W, hd = q.shape
    if k.shape != q.shape or tuple(flow_in.shape) != (B, 3, D, H, W):
        raise SmileError("modet_fused: q/k/flow_in shapes disagree")
    if rpb is not None:
        rpb = _chk(rpb, "rpb", 4)
        if tuple(rpb.shape) != (1, 3, 3, 3):
            raise SmileError("modet_fused handles heads == 1 only (rpb must be [1,3,3,3])")
    if (ln_gamma is None) != (ln_beta is None):
        raise SmileError("modet_fused: ln_gamma and ln_beta must be given together")
    if ln_gamma is not None:
        ln_gamma, ln_beta = _chk(ln_gamma, "ln_gamma", 1), _chk(ln_beta, "ln_beta", 1)
        if ln_gamma.numel() != hd or ln_beta.numel() != hd:
            raise SmileError(f"modet_fused: ln_gamma / ln_beta must have head_dim={hd} elements")
    flow_out = torch.empty_like(flow_in)
    moved = None
    cm = 0
    if moving is not None:
        moving = _chk(moving, "moving", 5)
        cm = moving.shape[1]
        if tuple(moving.shape) != (B, cm, D, H, W):
            raise SmileError("modet_fused: moving must be [B,C,D,H,W] at the flow resolution")
        moved = torch.empty_like(moving)
    call("smile_modet_fused_fwd", q.data_ptr(), k.data_ptr(), _ptr(rpb), _ptr(ln_gamma), _ptr(ln_beta), flow_in.data_ptr(), _ptr(moving),
         flow_out.data_ptr(), _ptr(moved), B, D, H, W, hd, float(scale), float(postmul), cm, _stream(),
         label=f"[{D}x{H}x{W} mov{cm}]")
    return flow_out, moved


def proj_ln(feat: Tensor, weight: Tensor, bias: Tensor, gamma: Tensor, beta: Tensor, eps: float = IN_EPS) -> Tensor:
    """ProjectionLayer.forward (ModeT/models.py:238-241): [B,Cin,D,H,W] -> [B,D,H,W,C]."""
    feat = _chk(feat, "feat", 5)
    weight = _chk(weight, "proj.weight", 2)
    B, Cin, D, H, W = feat.shape
    C = weight.shape[0]
    if weight.shape[1] != Cin:
        raise SmileError(f"proj.weight {tuple(weight.shape)} does not match Cin={Cin}")
    bias, gamma, beta = _chk(bias, "proj.bias", 1), _chk(gamma, "norm.weight", 1), _chk(beta, "norm.bias", 1)
    out = torch.empty((B, D, H, W, C), device=feat.device, dtype=torch.float32)
    call("smile_proj_ln_fwd", feat.data_ptr(), weight.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
         out.data_ptr(), B, Cin, C, D * H * W, float(eps), _stream(), label=f"[{Cin}->{C} {D}x{H}x{W}]")
    return out


FUSED_WARP_PROJ = {(8, 6), (16, 6), (32, 12), (64, 24)}   # (Cin, C) pairs compiled into smile_warp_proj_ln_fwd


def warp_proj_ln(src: Tensor, flow: Tensor, weight: Tensor, bias: Tensor, gamma: Tensor, beta: Tensor,
                 eps: float = IN_EPS) -> Tensor:
    """ProjectionLayer(SpatialTransformer(src, flow)) (ModeT/models.py:388-389 pattern) -> [B,D,H,W,C]; one kernel
    when (Cin, C) is one of the reference's decoder widths, otherwise warp3d followed by proj_ln."""
    src = _chk(src, "src", 5)
    flow = _chk(flow, "flow", 5)
    weight = _chk(weight, "proj.weight", 2)
    B, Cin, D, H, W = src.shape
    C = weight.shape[0]
    if (Cin, C) not in FUSED_WARP_PROJ or min(D, H, W) < 2:
        return proj_ln(warp3d(src, flow), weight, bias, gamma, beta, eps)
    if tuple(flow.shape) != (B, 3, D, H, W) or weight.shape[1] != Cin:
        raise SmileError("warp_proj_ln: src / flow / weight shapes disagree")
    bias, gamma, beta = _chk(bias, "proj.bias", 1), _chk(gamma, "norm.weight", 1), _chk(beta, "norm.bias", 1)
    out = torch.empty((B, D, H, W, C), device=src.device, dtype=torch.float32)
    call("smile_warp_proj_ln_fwd", src.data_ptr(), flow.data_ptr(), weight.data_ptr(), bias.data_ptr(), gamma.data_ptr(),
         beta.data_ptr(), out.data_ptr(), B, Cin, C, D, H, W, float(eps), _stream(), label=f"[{Cin}->{C} {D}x{H}x{W}]")
    return out


# Split / re-arranged weights for the tensor-core convolution, prepared once per nn.Parameter and refreshed when torch can
# see that the parameter changed: an in-place update through the autograd-tracked tensor (optimizer step, `p.copy_()`,
# `load_state_dict` -> version counter), a re-allocation (`.to()`, `.cuda()` -> data_ptr) or a new Parameter object.
# What torch canNOT see is a write through `p.data` (`p.data.copy_()`, `p.data.mul_()`: EMA / SWA loops, some custom
# initialisers) or through a raw pointer (our own fused Adam clears the cache itself): call
# `ops.invalidate_prepared_weights()` after such a write, or set `ops.PREPARED_WEIGHT_CACHE = False` to prepare the
# weights on every call (~1 % of a forward).  `ModeT.load_state_dict`, `.train()` and `.eval()` invalidate as well.
# Plain tensors (e.g. the flipped weights of the data-gradient pass) are never cached.
_WPREP_CACHE: dict = {}
PREPARED_WEIGHT_CACHE = True


def invalidate_prepared_weights() -> None:
    """Forget every prepared (split / re-arranged) tensor-core weight block; the next conv3d call prepares them again."""
    _WPREP_CACHE.clear()


def _prepared_weights(weight: Tensor, Cin: int, Cout: int) -> Optional[Tensor]:
    # inference only: a training step rewrites the parameters through raw pointers (fused Adam over the flat buffer),
    # which torch's version counter does not see
    if torch.is_grad_enabled() or not isinstance(weight, torch.nn.Parameter) or Cin < 16 or Cout < 12:
        return None
    import weakref
    from ._lib import lib
    key = id(weight)
    cur = torch.cuda.current_stream(weight.device)
    hit = _WPREP_CACHE.get(key) if PREPARED_WEIGHT_CACHE else None
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
        # prepared on another stream and possibly still in flight: order this stream after the preparation (a completed
        # event needs no edge -- and must not get one while this stream is being captured into a CUDA graph)
        # (graph.GraphedForward synchronises the device before it starts capturing, so every preparation is complete then)
        if hit[5] != cur.cuda_stream and not torch.cuda.is_current_stream_capturing() and not hit[4].query():
            cur.wait_event(hit[4])
        return hit[3]
    n = int(lib().smile_conv3d_tc_prep_floats(Cin, Cout))
    wprep = torch.empty(n, device=weight.device, dtype=torch.float32)
    call("smile_conv3d_tc_prep", weight.data_ptr(), wprep.data_ptr(), Cin, Cout, _stream())
    if PREPARED_WEIGHT_CACHE:
        done = torch.cuda.Event()
        done.record(cur)
        _WPREP_CACHE[key] = (weakref.ref(weight, lambda _r, k=key: _WPREP_CACHE.pop(k, None)), weight._version,
                             weight.data_ptr(), wprep, done, cur.cuda_stream)
    return wprep


# Precision of the MMA operands of conv3d: "fp32" (default: fp32 SIMT kernels + 3xTF32 tcgen05, fp32-class results) or
# "bf16" (tcgen05 kind::f16 with fp32 accumulation, csrc/conv_bf16.cu: BASELINE.json configs[2..3]).  Activations,
# statistics and every other operator stay fp32.  Set through the context manager below or `ModeT.conv_precision`.
_CONV_PRECISION = "fp32"


def current_conv_precision() -> str:
    return _CONV_PRECISION


class conv_precision:
    """with ops.conv_precision("bf16"): ...   -- conv3d (forward and the data-gradient pass) on bf16 tensor cores."""

    def __init__(self, mode: str):
        if mode not in ("fp32", "bf16"):
            raise ValueError(f"conv precision must be 'fp32' or 'bf16', got {mode!r}")
        self.mode = mode

    def __enter__(self):
        global _CONV_PRECISION
        self.prev, _CONV_PRECISION = _CONV_PRECISION, self.mode
        return self

    def __exit__(self, *exc):
        global _CONV_PRECISION
        _CONV_PRECISION = self.prev
        return False


def conv3d(x: Tensor, weight: Tensor, bias: Tensor, in_stats: Optional[Tensor] = None, want_stats: bool = False,
           act_out: bool = False, eps: float = IN_EPS) -> Tuple[Tensor, Optional[Tensor]]:
    """Conv3d(k=3, s=1, p=1) (ModeT/models.py:127, 143, 253).  With `in_stats` the input is a raw conv
    output whose InstanceNorm+LeakyReLU is applied on load; with `want_stats` the fp64 (sum, sumsq)
    of the raw output per (b, c) are returned for the next InstanceNorm."""
    x = _chk(x, "x", 5)
    weight_in = weight
    weight = _chk(weight, "weight", 5)
    bias = _chk(bias, "bias", 1)
    B, Cin, D, H, W = x.shape
    Cout = weight.shape[0]
    if tuple(weight.shape) != (Cout, Cin, 3, 3, 3):
        raise SmileError(f"weight must be [{Cout},{Cin},3,3,3], got {tuple(weight.shape)}")
    if in_stats is not None:
        in_stats = _chk(in_stats, "in_stats", 2, torch.float64)
        if tuple(in_stats.shape) != (B * Cin, 2):
            raise SmileError("in_stats must be [B*Cin, 2] float64")
    out = torch.empty((B, Cout, D, H, W), device=x.device, dtype=torch.float32)
    stats = stats_arena.take(B * Cout, x.device) if want_stats else None
    if _CONV_PRECISION == "bf16":
        call("smile_conv3d_bf16_fwd", x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(in_stats),
             _ptr(stats), B, Cin, Cout, D, H, W, int(act_out), float(eps), _stream(), label=f"[{Cin}->{Cout} {D}x{H}x{W}]")
        return out, stats
    wprep = _prepared_weights(weight_in, Cin, Cout) if weight is weight_in else None
    if wprep is not None:
        call("smile_conv3d_prepped_fwd", x.data_ptr(), weight.data_ptr(), wprep.data_ptr(), bias.data_ptr(), out.data_ptr(),
             _ptr(in_stats), _ptr(stats), B, Cin, Cout, D, H, W, int(act_out), float(eps), _stream(),
             label=f"[{Cin}->{Cout} {D}x{H}x{W}]")
        # the breakdown tools key on the smile_conv3d_fwd label
    else:
        call("smile_conv3d_fwd", x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(in_stats),
             _ptr(stats), B, Cin, Cout, D, H, W, int(act_out), float(eps), _stream(), label=f"[{Cin}->{Cout} {D}x{H}x{W}]")
    return out, stats


def instnorm_lrelu_pool(raw: Tensor, stats: Tensor, pool: bool = False, inplace: bool = False,
                        eps: float = IN_EPS) -> Tuple[Tensor, Optional[Tensor]]:
    """InstanceNorm3d + LeakyReLU(0.1) (ModeT/models.py:149-150) from fp64 sums; optionally also
    AvgPool3d(2) of the result (models.py:198)."""
    raw = _chk(raw, "raw", 5)
    stats = _chk(stats, "stats", 2, torch.float64)
    B, C, D, H, W = raw.shape
    out = raw if inplace else torch.empty_like(raw)
    pooled = torch.empty((B, C, D // 2, H // 2, W // 2), device=raw.device, dtype=torch.float32) if pool else None
    call("smile_instnorm_lrelu_pool_fwd", raw.data_ptr(), stats.data_ptr(), out.data_ptr(), _ptr(pooled), B, C, D, H, W,
         float(eps), _stream(), label=f"[c{C} {D}x{H}x{W}]")
    return out, pooled


def cwm_fuse(fields: Tensor, logits: Tensor) -> Tensor:
    """CWM tail (ModeT/models.py:254, 268-275): 2 * sum_f fields[:, 3f:3f+3] * softmax(logits)[:, f]."""
    fields = _chk(fields, "fields", 5)
    logits = _chk(logits, "logits", 5)
    B, F3, D, H, W = fields.shape
    F = logits.shape[1]
    if F3 != 3 * F or tuple(logits.shape) != (B, F, D, H, W):
        raise SmileError(f"fields {tuple(fields.shape)} / logits {tuple(logits.shape)} mismatch")
    out = torch.empty((B, 3, D, H, W), device=fields.device, dtype=torch.float32)
    call("smile_cwm_fuse_fwd", fields.data_ptr(), logits.data_ptr(), out.data_ptr(), B, F, D * H * W, _stream(),
         label=f"[f{F} {D}x{H}x{W}]")
    return out


def modet_qkrpb_fwd(query: Tensor, key: Tensor, rpb: Optional[Tensor]) -> Tensor:
    """Twin of the reference's `modet.modet_fw` (ModeT-cu/modet/modet.cpp:4-18): pre-softmax logits.
    query [B,h,H,W,T,d], key [B,h,H+2,W+2,T+2,d] (zero padded), rpb [h,3,3,3] or None -> [B,h,H,W,T,27]."""
    query = _chk(query, "query", 6)
    key = _chk(key, "key", 6)
    B, h, H, W, T, d = query.shape
    if tuple(key.shape) != (B, h, H + 2, W + 2, T + 2, d):
        raise SmileError(f"key must be the zero-padded [{B},{h},{H + 2},{W + 2},{T + 2},{d}], got {tuple(key.shape)}")
    if rpb is not None:
        rpb = _chk(rpb, "rpb", 4)
        if tuple(rpb.shape) != (h, 3, 3, 3):
            raise SmileError(f"rpb must be [{h},3,3,3], got {tuple(rpb.shape)}")
    attn = torch.empty((B, h, H, W, T, 27), device=query.device, dtype=torch.float32)
    call("smile_modet_qkrpb_fwd", query.data_ptr(), key.data_ptr(), _ptr(rpb), attn.data_ptr(), B, h, H, W, T, d, _stream(),
         label=f"[h{h} {H}x{W}x{T}]")
    return attn


def modet_qkrpb_bwd(d_attn: Tensor, query: Tensor, key: Tensor, bias: bool):
    """Twin of `modet.modet_bw` (modet.cpp:20-31): returns (d_query, d_key_padded, d_rpb or None)."""
    d_attn = _chk(d_attn, "d_attn", 6)
    query = _chk(query, "query", 6)
    key = _chk(key, "key", 6)
    B, h, H, W, T, d = query.shape
    if tuple(d_attn.shape) != (B, h, H, W, T, 27) or tuple(key.shape) != (B, h, H + 2, W + 2, T + 2, d):
        raise SmileError("modet_qkrpb_bwd: d_attn / query / key shapes disagree")
    dq = torch.empty_like(query)
    dk = torch.empty_like(key)
    drpb = torch.empty((h, 3, 3, 3), device=query.device, dtype=torch.float32) if bias else None
    call("smile_modet_qkrpb_bwd", d_attn.data_ptr(), query.data_ptr(), key.data_ptr(), dq.data_ptr(), dk.data_ptr(),
         _ptr(drpb), B, h, H, W, T, d, _stream(), label=f"[h{h} {H}x{W}x{T}]")
    return dq, dk, drpb


def ncc_vxm(y_true: Tensor, y_pred: Tensor, win: int = 9) -> Tensor:
    """NCC_vxm.forward (ModeT/losses.py:43-95): -mean local normalised cross-correlation, scalar tensor."""
    y_true = _chk(y_true, "y_true", 5)
    y_pred = _chk(y_pred, "y_pred", 5)
    B, C, D, H, W = y_true.shape
    if C != 1 or y_pred.shape != y_true.shape:
        raise SmileError("ncc_vxm: y_true and y_pred must both be [B,1,D,H,W]")
    from ._lib import lib
    nbytes = lib().smile_ncc_vxm_work_bytes(B, D, H, W)
    work = torch.empty(nbytes, device=y_true.device, dtype=torch.uint8)
    out = torch.empty(1, device=y_true.device, dtype=torch.float32)
    call("smile_ncc_vxm_fwd", y_true.data_ptr(), y_pred.data_ptr(), out.data_ptr(), work.data_ptr(), B, D, H, W, int(win),
         _stream(), label=f"[{D}x{H}x{W}]")
    return out[0]


def grad3d_l2(flow: Tensor) -> Tensor:
    """Grad3d(penalty='l2').forward (ModeT/losses.py:16-31), scalar tensor."""
    flow = _chk(flow, "flow", 5)
    B, C, D, H, W = flow.shape
    work = torch.empty(4, device=flow.device, dtype=torch.float64)
    out = torch.empty(1, device=flow.device, dtype=torch.float32)
    call("smile_grad3d_l2_fwd", flow.data_ptr(), out.data_ptr(), work.data_ptr(), B, C, D, H, W, _stream(),
         label=f"[{D}x{H}x{W}]")
    return out[0]


# ------------------------------------------------------------------------------------------------
# backward operators (training path); thin wrappers like the forward ones
# ------------------------------------------------------------------------------------------------
def warp3d_bwd(g: Tensor, src: Tensor, flow: Tensor, need_src: bool = True, need_flow: bool = True):
    """Gradients of warp3d w.r.t. src and flow (SpatialTransformer backward)."""
    g, src, flow = _chk(g, "g", 5), _chk(src, "src", 5), _chk(flow, "flow", 5)
    B, C, D, H, W = src.shape
    d_src = torch.empty_like(src) if need_src else None
    d_flow = torch.empty_like(flow) if need_flow else None
    call("smile_warp3d_bwd", g.data_ptr(), src.data_ptr(), flow.data_ptr(), _ptr(d_src), _ptr(d_flow), B, C, D, H, W,
         _stream(), label=f"[c{C} {D}x{H}x{W}]")
    return d_src, d_flow


def upsample2x_bwd(g: Tensor, premul: float = 1.0) -> Tensor:
    g = _chk(g, "g", 5)
    B, C, OD, OH, OW = g.shape
    dx = torch.empty((B, C, OD // 2, OH // 2, OW // 2), device=g.device, dtype=torch.float32)
    call("smile_upsample2x_bwd", g.data_ptr(), dx.data_ptr(), B, C, OD // 2, OH // 2, OW // 2, float(premul), _stream(),
         label=f"[c{C} {OD // 2}x{OH // 2}x{OW // 2}]")
    return dx


def modet_attention_bwd(g: Tensor, q: Tensor, k: Tensor, rpb: Optional[Tensor], heads: int, scale: float):
    g, q, k = _chk(g, "g", 5), _chk(q, "q", 5), _chk(k, "k", 5)
    B, D, H, W, C = q.shape
    if rpb is not None:
        rpb = _chk(rpb, "rpb", 4)
    dq, dk = torch.empty_like(q), torch.empty_like(k)
    drpb = torch.empty((heads, 3, 3, 3), device=q.device, dtype=torch.float32) if rpb is not None else None
    work = torch.empty(B * D * H * W * heads * 27, device=q.device, dtype=torch.float32)
    call("smile_modet_attn_bwd", g.data_ptr(), q.data_ptr(), k.data_ptr(), _ptr(rpb), dq.data_ptr(), dk.data_ptr(),
         _ptr(drpb), work.data_ptr(), B, D, H, W, heads, C // heads, float(scale), _stream(), label=f"[h{heads} {D}x{H}x{W}]")
    return dq, dk, drpb


def proj_ln_bwd(g: Tensor, feat: Tensor, weight: Tensor, bias: Tensor, gamma: Tensor, eps: float = IN_EPS,
                need_feat: bool = True):
    g, feat, weight = _chk(g, "g", 5), _chk(feat, "feat", 5), _chk(weight, "proj.weight", 2)
    bias, gamma = _chk(bias, "proj.bias", 1), _chk(gamma, "norm.weight", 1)
    B, Cin, D, H, W = feat.shape
    C = weight.shape[0]
    dfeat = torch.empty_like(feat) if need_feat else None
    dw, db = torch.empty_like(weight), torch.empty_like(bias)
    dg, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
    call("smile_proj_ln_bwd", g.data_ptr(), feat.data_ptr(), weight.data_ptr(), bias.data_ptr(), gamma.data_ptr(),
         _ptr(dfeat), dw.data_ptr(), db.data_ptr(), dg.data_ptr(), dbeta.data_ptr(), B, Cin, C, D * H * W, float(eps),
         _stream(), label=f"[{Cin}->{C} {D}x{H}x{W}]")
    return dfeat, dw, db, dg, dbeta


def cwm_fuse_bwd(g: Tensor, fields: Tensor, logits: Tensor):
    g, fields, logits = _chk(g, "g", 5), _chk(fields, "fields", 5), _chk(logits, "logits", 5)
    B, F = logits.shape[0], logits.shape[1]
    N = logits.shape[2] * logits.shape[3] * logits.shape[4]
    dfields, dlogits = torch.empty_like(fields), torch.empty_like(logits)
    call("smile_cwm_fuse_bwd", g.data_ptr(), fields.data_ptr(), logits.data_ptr(), dfields.data_ptr(), dlogits.data_ptr(),
         B, F, N, _stream(), label=f"[f{F}]")
    return dfields, dlogits


def conv3d_bwd(g: Tensor, x: Tensor, weight: Tensor, need_x: bool = True, need_bias: bool = True):
    """Gradients of conv3d (k=3, s=1, p=1): d_x through the forward kernel with flipped weights, d_weight, d_bias."""
    g, x, weight = _chk(g, "g", 5), _chk(x, "x", 5), _chk(weight, "weight", 5)
    B, Cin, D, H, W = x.shape
    Cout = weight.shape[0]
    dx = None
    if need_x:
        wT = torch.empty((Cin, Cout, 3, 3, 3), device=x.device, dtype=torch.float32)
        call("smile_conv3d_flip_weights", weight.data_ptr(), wT.data_ptr(), Cout, Cin, _stream())
        zero_b = torch.zeros(Cin, device=x.device, dtype=torch.float32)
        dx, _ = conv3d(g, wT, zero_b)
    dw = torch.empty_like(weight)
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if need_bias else None
    # bf16 mode: the weight-gradient products also run on the tensor cores (wgrad_tc.cu)
    entry = "smile_conv3d_wgrad_bf16" if _CONV_PRECISION == "bf16" else "smile_conv3d_wgrad"
    call(entry, x.data_ptr(), g.data_ptr(), dw.data_ptr(), _ptr(db), B, Cin, Cout, D, H, W, _stream(),
         label=f"[{Cin}->{Cout} {D}x{H}x{W}]")
    return dx, dw, db


def in_lrelu_bwd(d_act: Tensor, act: Tensor, stats: Optional[Tensor], mode: int = 0, eps: float = IN_EPS) -> Tensor:
    """d_raw of act = LeakyReLU(InstanceNorm(raw)) (mode 0) or act = LeakyReLU(raw) (mode 1)."""
    d_act, act = _chk(d_act, "d_act", 5), _chk(act, "act", 5)
    B, C, D, H, W = act.shape
    work = torch.empty((B * C, 2), device=act.device, dtype=torch.float64) if mode == 0 else None
    if mode == 0:
        stats = _chk(stats, "stats", 2, torch.float64)
    d_raw = torch.empty_like(act)
    call("smile_in_lrelu_bwd", d_act.data_ptr(), act.data_ptr(), _ptr(stats) if mode == 0 else None, _ptr(work),
         d_raw.data_ptr(), B, C, D * H * W, float(eps), int(mode), _stream(), label=f"[c{C} {D}x{H}x{W}]")
    return d_raw


def avgpool2_bwd_add(d_pooled: Tensor, d_full: Tensor) -> Tensor:
    """d_full += AvgPool3d(2) backward of d_pooled (in place)."""
    d_pooled, d_full = _chk(d_pooled, "d_pooled", 5), _chk(d_full, "d_full", 5)
    B, C, D, H, W = d_full.shape
    call("smile_avgpool2_bwd_add", d_pooled.data_ptr(), d_full.data_ptr(), B, C, D, H, W, _stream())
    return d_full


def ncc_vxm_bwd(y_true: Tensor, y_pred: Tensor, gscale: Optional[Tensor], win: int = 9) -> Tensor:
    y_true, y_pred = _chk(y_true, "y_true", 5), _chk(y_pred, "y_pred", 5)
    B, C, D, H, W = y_true.shape
    from ._lib import lib
    work = torch.empty(lib().smile_ncc_vxm_work_bytes(B, D, H, W), device=y_true.device, dtype=torch.uint8)
    d_true = torch.empty_like(y_true)
    call("smile_ncc_vxm_bwd", y_true.data_ptr(), y_pred.data_ptr(), d_true.data_ptr(), work.data_ptr(), _ptr(gscale), B, D,
         H, W, int(win), _stream())
    return d_true


def grad3d_l2_bwd(flow: Tensor, gscale: Optional[Tensor]) -> Tensor:
    flow = _chk(flow, "flow", 5)
    B, C, D, H, W = flow.shape
    d_flow = torch.empty_like(flow)
    call("smile_grad3d_l2_bwd", flow.data_ptr(), d_flow.data_ptr(), _ptr(gscale), B, C, D, H, W, _stream())
    return d_flow


def adam_amsgrad_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, max_exp_avg_sq: Tensor, lr: float,
                      beta1: float, beta2: float, eps: float, step: int) -> None:
    """In-place torch.optim.Adam(amsgrad=True) update of a flat fp32 buffer."""
    _WPREP_CACHE.clear()   # the parameters change behind torch's version counters
    for t in (param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise SmileError("adam_amsgrad_step: all buffers must be contiguous fp32 CUDA tensors")
    call("smile_adam_amsgrad_step", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
         max_exp_avg_sq.data_ptr(), param.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step), _stream())
