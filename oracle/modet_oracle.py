"""CPU oracle for the ModeT registration hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in elementary torch-CPU tensor arithmetic, of what the
reference's `ModeT/models.py` computes on the hot path.  It is the
*checker*: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  Nothing under `smilecode_b200/` imports it and the
product path raises when the CUDA library is missing instead of routing here.

Parity pin: the reference ships no tests / golden vectors of its own,
and all of its arithmetic is delegated to PyTorch (un-vendored, unpinned; installed here:
torch 2.11.0).  The pin is therefore `tests/golden/*.npz`, produced by importing the
reference's own `ModeT/models.py` in the build container (`oracle/make_golden.py`,
committed) -- `tests/test_oracle_golden.py` checks every function below against them.

Every function cites the reference lines it follows (paths under /root/reference).
Formulas for the two library ops whose integer indices must be bit exact come from the
torch headers shipped in the wheel: ATen/native/GridSampler.h:27-36 (un-normalise) and
ATen/native/UpSample.h:271-296,451-475 (align_corners scale, index / lambda).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# a2  ModeTransformer.forward   (ModeT/models.py:308-334, makeV 298-301, apply_pb 303-306)
# --------------------------------------------------------------------------------------
def modet_attention(q: Tensor, k: Tensor, rpb: Tensor | None, heads: int, scale: float) -> Tensor:
    """q, k: [B, D, H, W, heads*hd] channels-last.  Returns [B, 3*heads, D, H, W].

    logit[t] = scale * <q[n, h], kpad[n + off(t), h]> + rpb[h, t]   over the 27 taps of the
    zero-padded key volume (models.py:319: the padded taps stay inside the softmax with
    logit = rpb only); out[h*3 + a] = sum_t softmax(logit)[t] * off(t)[a]  (models.py:328-332,
    V = grid buffer built at 293-296: tap t=(i*3+j)*3+k has V[t] = (i-1, j-1, k-1)).
    """
    B, D, H, W, C = q.shape
    hd = C // heads
    qh = (q.reshape(B, D, H, W, heads, hd) * scale)
    kp = F.pad(k.reshape(B, D, H, W, heads, hd), (0, 0, 0, 0, 1, 1, 1, 1, 1, 1))
    logits = []
    for t in range(27):
        ti, tj, tk = t // 9, (t // 3) % 3, t % 3
        kt = kp[:, ti:ti + D, tj:tj + H, tk:tk + W]
        lg = (qh * kt).sum(-1)  # [B, D, H, W, heads]
        if rpb is not None:
            lg = lg + rpb[:, ti, tj, tk]
        logits.append(lg)
    p = torch.stack(logits, -1).softmax(-1)  # [B, D, H, W, heads, 27]
    off = torch.tensor([[t // 9 - 1, (t // 3) % 3 - 1, t % 3 - 1] for t in range(27)],
                       dtype=q.dtype, device=q.device)
    out = p @ off  # [B, D, H, W, heads, 3]
    return out.reshape(B, D, H, W, heads * 3).permute(0, 4, 1, 2, 3).contiguous()


# --------------------------------------------------------------------------------------
# a5  SpatialTransformer.forward   (ModeT/models.py:49-67) + torch grid_sample semantics
# --------------------------------------------------------------------------------------
def warp_source_coords(flow: Tensor) -> Tensor:
    """Per-axis un-normalised sampling coordinate, replaying the fp32 op sequence exactly:
    p = idx + flow (models.py:51); n = 2*(p/(S-1) - 0.5) (models.py:56);
    x = ((n + 1)/2)*(S-1) (GridSampler.h:31).  flow [B,3,D,H,W] -> x [B,3,D,H,W]."""
    B, three, D, H, W = flow.shape
    shape = (D, H, W)
    xs = []
    for a, S in enumerate(shape):
        view = [1, 1, 1]
        view[a] = S
        idx = torch.arange(S, dtype=flow.dtype, device=flow.device).view(1, *view)
        p = idx + flow[:, a]
        n = 2 * (p / (S - 1) - 0.5)
        xs.append(((n + 1) / 2) * (S - 1))
    return torch.stack(xs, 1)


def warp_corner_indices(flow: Tensor) -> Tensor:
    """Integer floor() corner indices [B,3,D,H,W] int64 -- the bit-exact part of a5."""
    return torch.floor(warp_source_coords(flow)).to(torch.int64)


def warp_trilinear(src: Tensor, flow: Tensor) -> Tensor:
    """src [B,C,D,H,W], flow [B,3,D,H,W] (channel a = displacement along dim a, in voxels).
    Trilinear, zeros padding, align_corners=True; eight corners accumulated in torch's order
    tnw,tne,tsw,tse,bnw,bne,bsw,bse with weights (x1-x)*(y1-y)*(z1-z) etc."""
    B, C, D, H, W = src.shape
    x = warp_source_coords(flow)
    z, y, xw = x[:, 0], x[:, 1], x[:, 2]          # D-, H-, W-axis coordinates
    z0f, y0f, x0f = torch.floor(z), torch.floor(y), torch.floor(xw)
    z1f, y1f, x1f = z0f + 1, y0f + 1, x0f + 1
    wz = (z1f - z, z - z0f)
    wy = (y1f - y, y - y0f)
    wx = (x1f - xw, xw - x0f)
    zi = (z0f.long(), z0f.long() + 1)
    yi = (y0f.long(), y0f.long() + 1)
    xi = (x0f.long(), x0f.long() + 1)
    flat = src.reshape(B, C, D * H * W)
    out = torch.zeros_like(src)
    for cz in (0, 1):            # t / b
        for cy in (0, 1):        # n / s
            for cx in (0, 1):    # w / e
                w = (wx[cx] * wy[cy]) * wz[cz]
                inb = ((zi[cz] >= 0) & (zi[cz] < D) & (yi[cy] >= 0) & (yi[cy] < H) &
                       (xi[cx] >= 0) & (xi[cx] < W))
                lin = (zi[cz].clamp(0, D - 1) * H + yi[cy].clamp(0, H - 1)) * W + xi[cx].clamp(0, W - 1)
                g = torch.gather(flat, 2, lin.reshape(B, 1, -1).expand(B, C, -1)).reshape(src.shape)
                contrib = g * w.unsqueeze(1)
                out = out + torch.where(inb.unsqueeze(1), contrib, torch.zeros_like(contrib))
    return out


# --------------------------------------------------------------------------------------
# a6  nn.Upsample(scale_factor=2, trilinear, align_corners=True)   (ModeT/models.py:354, 257-261)
# --------------------------------------------------------------------------------------
def upsample2x_indices(S: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(i0, i1, lambda1) for every output index of one axis of size S -> 2S
    (UpSample.h:277-296 scale, 451-475 index/lambda; all fp32)."""
    O = 2 * S
    ratio = torch.tensor(float(S - 1), dtype=torch.float32) / torch.tensor(float(O - 1), dtype=torch.float32) \
        if O > 1 else torch.tensor(0.0)
    dst = torch.arange(O, dtype=torch.float32)
    real = ratio * dst
    i0 = real.to(torch.int64).clamp(max=S - 1)
    lam1 = (real - i0.to(torch.float32)).clamp(0, 1)
    i1 = i0 + (i0 < S - 1).to(torch.int64)
    return i0, i1, lam1


def upsample2x_trilinear(x: Tensor) -> Tensor:
    """x [B,C,D,H,W] -> [B,C,2D,2H,2W], nested-lerp form (W innermost, then H, then D)."""
    B, C, D, H, W = x.shape
    d0, d1, ld = [t.to(x.device) for t in upsample2x_indices(D)]
    h0, h1, lh = [t.to(x.device) for t in upsample2x_indices(H)]
    w0, w1, lw = [t.to(x.device) for t in upsample2x_indices(W)]
    ld = ld.view(1, 1, -1, 1, 1)
    lh = lh.view(1, 1, 1, -1, 1)
    lw = lw.view(1, 1, 1, 1, -1)

    def lerp_w(v):   # v [B,C,d,h,W] -> [B,C,d,h,2W]
        return (1 - lw) * v[..., w0] + lw * v[..., w1]

    def lerp_hw(v):  # v [B,C,d,H,W]
        return (1 - lh) * lerp_w(v[:, :, :, h0]) + lh * lerp_w(v[:, :, :, h1])

    return (1 - ld) * lerp_hw(x[:, :, d0]) + ld * lerp_hw(x[:, :, d1])


# --------------------------------------------------------------------------------------
# building blocks used by a4 / a8: Conv3d 3x3x3 pad 1, InstanceNorm3d, LeakyReLU(0.1)
# (ModeT/models.py:119-151).  torch's fp32 CPU conv is the floating-point reference.
# --------------------------------------------------------------------------------------
def conv3(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.conv3d(x, w, b, stride=1, padding=1)


def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """InstanceNorm3d(affine=False, track_running_stats=False): biased variance over D*H*W."""
    m = x.mean(dim=(2, 3, 4), keepdim=True)
    v = x.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
    return (x - m) / torch.sqrt(v + eps)


def lrelu(x: Tensor) -> Tensor:
    return torch.where(x >= 0, x, 0.1 * x)


def conv_ins_block(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    return lrelu(instance_norm(conv3(x, sd[prefix + ".main.weight"], sd[prefix + ".main.bias"])))


# --------------------------------------------------------------------------------------
# a4  CWM.forward   (ModeT/models.py:263-275; layers 250-261)
# --------------------------------------------------------------------------------------
def cwm(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """x [B,3F,D,H,W] -> [B,3,2D,2H,2W]:  u = up2(x); softmax-over-F weights from three convs;
    out = 2 * sum_f u[:, 3f:3f+3] * wgt[:, f]."""
    u = upsample2x_trilinear(x)
    t = conv_ins_block(u, sd, prefix + ".conv.0")
    t = conv_ins_block(t, sd, prefix + ".conv.1")
    wgt = conv3(t, sd[prefix + ".conv.2.weight"], sd[prefix + ".conv.2.bias"]).softmax(dim=1)
    Fn = x.shape[1] // 3
    acc = 0
    for f in range(Fn):
        acc = acc + u[:, 3 * f:3 * f + 3] * wgt[:, f:f + 1]
    return 2 * acc


# --------------------------------------------------------------------------------------
# a7  ProjectionLayer.forward   (ModeT/models.py:238-241)
# --------------------------------------------------------------------------------------
def projection(feat: Tensor, sd: Dict[str, Tensor], prefix: str, eps: float = 1e-5) -> Tensor:
    """feat [B,Cin,D,H,W] -> LayerNorm(Linear(feat channels-last)) [B,D,H,W,C]."""
    y = feat.permute(0, 2, 3, 4, 1) @ sd[prefix + ".proj.weight"].t() + sd[prefix + ".proj.bias"]
    m = y.mean(-1, keepdim=True)
    v = y.var(-1, unbiased=False, keepdim=True)
    return (y - m) / torch.sqrt(v + eps) * sd[prefix + ".norm.weight"] + sd[prefix + ".norm.bias"]


# --------------------------------------------------------------------------------------
# a8  Encoder.forward   (ModeT/models.py:186-228)
# --------------------------------------------------------------------------------------
def encoder(x: Tensor, sd: Dict[str, Tensor], prefix: str = "encoder") -> List[Tensor]:
    p = prefix
    t = lrelu(conv3(x, sd[p + ".conv0.0.main.weight"], sd[p + ".conv0.0.main.bias"]))
    t = conv_ins_block(t, sd, p + ".conv0.1")
    outs = [conv_ins_block(t, sd, p + ".conv0.2")]
    for lvl in range(1, 5):
        t = F.avg_pool3d(outs[-1], 2)
        t = conv_ins_block(t, sd, f"{p}.conv{lvl}.1")
        outs.append(conv_ins_block(t, sd, f"{p}.conv{lvl}.2"))
    return outs


# --------------------------------------------------------------------------------------
# a9  ModeT.forward   (ModeT/models.py:377-412)
# --------------------------------------------------------------------------------------
def _lib_warp(src: Tensor, flow: Tensor) -> Tensor:
    """SpatialTransformer exactly as the reference spells it (models.py:51-67): torch's own
    grid_sample does the sampling.  Used for the CPU-baseline timing leg."""
    shape = flow.shape[2:]
    idx = torch.stack(torch.meshgrid(*[torch.arange(0, s, dtype=flow.dtype) for s in shape], indexing="ij")).unsqueeze(0)
    loc = idx + flow
    for i, s in enumerate(shape):
        loc[:, i] = 2 * (loc[:, i] / (s - 1) - 0.5)
    return F.grid_sample(src, loc.permute(0, 2, 3, 4, 1)[..., [2, 1, 0]], align_corners=True, mode="bilinear")


def _lib_up2(x: Tensor) -> Tensor:
    return F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)   # models.py:354


def modet_forward(moving: Tensor, fixed: Tensor, sd: Dict[str, Tensor], num_heads: Sequence[int] = (8, 4, 2, 1, 1),
                  scale: float | None = 1.0, head_dim: int = 6, library_ops: bool = False) -> Tuple[Tensor, Tensor]:
    """Coarse-to-fine pyramid.  `scale=None` means head_dim**-0.5 (models.py:285).
    library_ops=True swaps the elementary warp / upsample restatements for the torch library calls
    the reference itself makes (same results, see tests) -- the faster, reference-like CPU timing."""
    sc = scale if scale else head_dim ** -0.5
    warp_trilinear = _lib_warp if library_ops else globals()["warp_trilinear"]
    upsample2x_trilinear = _lib_up2 if library_ops else globals()["upsample2x_trilinear"]
    M = encoder(moving, sd)
    Fx = encoder(fixed, sd)

    def attn(level: int, feat_m: Tensor) -> Tensor:
        heads = num_heads[5 - level]
        q = projection(Fx[level - 1], sd, f"projblock{level}")
        k = projection(feat_m, sd, f"projblock{level}")
        return modet_attention(q, k, sd[f"mdt{level}.rpb"], heads, sc)

    flow = cwm(attn(5, M[4]), sd, "cwm5")                                   # 383-386
    w = cwm(attn(4, warp_trilinear(M[3], flow)), sd, "cwm4")               # 388-391  (cwm keeps the restated upsample)
    up = upsample2x_trilinear(2 * flow)
    flow = warp_trilinear(up, w) + w                                        # 392
    w = cwm(attn(3, warp_trilinear(M[2], flow)), sd, "cwm3")               # 394-397
    up = upsample2x_trilinear(2 * flow)
    flow = warp_trilinear(up, w) + w                                        # 398
    w = attn(2, warp_trilinear(M[1], flow))                                 # 400-402
    flow = upsample2x_trilinear(2 * (warp_trilinear(flow, w) + w))          # 403
    w = attn(1, warp_trilinear(M[0], flow))                                 # 405-407
    flow = warp_trilinear(flow, w) + w                                      # 408
    return warp_trilinear(moving, flow), flow                               # 410-412


# --------------------------------------------------------------------------------------
# a10  losses  (ModeT/losses.py:43-95 NCC_vxm, 16-31 Grad3d 'l2'), device-agnostic restatement
# --------------------------------------------------------------------------------------
def ncc_vxm(y_true: Tensor, y_pred: Tensor, win: int = 9) -> Tensor:
    Ii, Ji = y_true, y_pred
    filt = torch.ones(1, 1, win, win, win, dtype=Ii.dtype, device=Ii.device)
    pad = win // 2
    box = lambda t: F.conv3d(t, filt, stride=1, padding=pad)
    I_sum, J_sum, I2_sum, J2_sum, IJ_sum = box(Ii), box(Ji), box(Ii * Ii), box(Ji * Ji), box(Ii * Ji)
    n = float(win ** 3)
    u_I, u_J = I_sum / n, J_sum / n
    cross = IJ_sum - u_J * I_sum - u_I * J_sum + u_I * u_J * n
    I_var = I2_sum - 2 * u_I * I_sum + u_I * u_I * n
    J_var = J2_sum - 2 * u_J * J_sum + u_J * u_J * n
    cc = cross * cross / (I_var * J_var + 1e-5)
    return -cc.mean()


def grad3d_l2(flow: Tensor) -> Tensor:
    dy = flow[:, :, 1:] - flow[:, :, :-1]
    dx = flow[:, :, :, 1:] - flow[:, :, :, :-1]
    dz = flow[:, :, :, :, 1:] - flow[:, :, :, :, :-1]
    return ((dx * dx).mean() + (dy * dy).mean() + (dz * dz).mean()) / 3.0


# --------------------------------------------------------------------------------------
# deterministic parameters for parity runs
# --------------------------------------------------------------------------------------
def param_shapes(channels: int = 4, head_dim: int = 6, num_heads: Sequence[int] = (8, 4, 2, 1, 1),
                 in_channel: int = 1) -> Dict[str, Tuple[int, ...]]:
    """Learnable tensors of ModeT in reference state_dict naming (models.py:186-219, 230-236,
    250-254, 292, 351-371)."""
    c = channels
    shp: Dict[str, Tuple[int, ...]] = {}

    def conv(name, cin, cout):
        shp[name + ".weight"] = (cout, cin, 3, 3, 3)
        shp[name + ".bias"] = (cout,)

    conv("encoder.conv0.0.main", in_channel, c)
    conv("encoder.conv0.1.main", c, 2 * c)
    conv("encoder.conv0.2.main", 2 * c, 2 * c)
    for lvl in range(1, 5):
        cin, cout = (2 ** lvl) * c, (2 ** (lvl + 1)) * c
        conv(f"encoder.conv{lvl}.1.main", cin, cout)
        conv(f"encoder.conv{lvl}.2.main", cout, cout)
    for level in range(1, 6):
        heads = num_heads[5 - level]
        dim = head_dim * heads
        cin = (2 ** level) * c
        shp[f"projblock{level}.proj.weight"] = (dim, cin)
        shp[f"projblock{level}.proj.bias"] = (dim,)
        shp[f"projblock{level}.norm.weight"] = (dim,)
        shp[f"projblock{level}.norm.bias"] = (dim,)
        shp[f"mdt{level}.rpb"] = (heads, 3, 3, 3)
        if level >= 3:
            fin, ch = 3 * heads, 6 * heads
            conv(f"cwm{level}.conv.0.main", fin, ch)
            conv(f"cwm{level}.conv.1.main", ch, ch)
            conv(f"cwm{level}.conv.2", ch, heads)
    return shp


def synth_state_dict(seed: int = 1234, ln_gain: Tuple[float, float] = (0.25, 0.75), rpb_std: float = 0.5,
                     **kw) -> Dict[str, Tensor]:
    """Seeded, non-degenerate weights: conv weights/biases U(+-1/sqrt(fan_in)) (the nn.Conv3d
    default bound), proj.weight ~ N(0, 1/Cin), proj.bias 0, LayerNorm weight ~ U(ln_gain),
    bias ~ N(0, 0.05), rpb ~ N(0, rpb_std).  Keys are drawn in sorted order from one CPU
    generator so the same dict is rebuilt on any box with the same torch."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in sorted(param_shapes(**kw).items()):
        if name.endswith("rpb"):
            t = torch.randn(shape, generator=g) * rpb_std
        elif ".proj.weight" in name:
            t = torch.randn(shape, generator=g) / math.sqrt(shape[1])
        elif ".proj.bias" in name:
            t = torch.zeros(shape)
        elif ".norm.weight" in name:
            t = torch.rand(shape, generator=g) * (ln_gain[1] - ln_gain[0]) + ln_gain[0]
        elif ".norm.bias" in name:
            t = torch.randn(shape, generator=g) * 0.05
        else:  # conv weight / bias
            wshape = shape if len(shape) == 5 else param_shapes(**kw)[name[:-4] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1] * 27)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t
    return sd
