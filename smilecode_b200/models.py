"""Drop-in replacement for the reference's `ModeT/models.py` (and `ModeT-cu/models.py`).

Same class names, constructor signatures, child-module names and state_dict keys as the reference
(SURVEY.md section 8b), so `from models import ModeT` in the reference's train.py / infer.py can be
pointed here (see INTEGRATION.md and dropin/models.py) and reference checkpoints load with
strict=True.  Every forward below runs hand-written sm_100a kernels through the C ABI
(`smilecode_b200.ops`); the nn.Conv3d / nn.Linear / nn.LayerNorm children are parameter holders
only (they give the reference's default initialisation and key names) and are never called.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as nnf
from torch.distributions.normal import Normal

from . import ops

# Debug / bisect switch (tools/parity_bisect.py): False runs levels 2 and 1 as the reference spells them
# (attention -> SpatialTransformer(flow, w) + w -> upsample / final warp: the stand-alone bit-exact kernels)
# instead of the fused single-head kernel.
FUSE_LEVELS = True

__all__ = ["SpatialTransformer", "VecInt", "ResizeTransform", "ConvBlock", "ConvInsBlock", "UpConvBlock",
           "DeconvBlock", "Encoder", "ProjectionLayer", "CWM", "ModeTransformer", "ModeT", "ModeT_cu"]


def _voxel_grid(size: Sequence[int]) -> torch.Tensor:
    axes = [torch.arange(0, int(s), dtype=torch.float32) for s in size]
    return torch.stack(torch.meshgrid(*axes, indexing="ij")).unsqueeze(0)


class SpatialTransformer(nn.Module):
    """Reference ModeT/models.py:25-67.  The identity `grid` buffer is kept only because it is part
    of the reference's state_dict; the kernel derives voxel coordinates from thread indices."""

    def __init__(self, size, mode: str = "bilinear"):
        super().__init__()
        self.mode = mode
        self.register_buffer("grid", _voxel_grid(size))

    def forward(self, src, flow):
        if flow.dim() != 5:
            raise NotImplementedError("SpatialTransformer: only 3-D volumes are supported")
        if self.mode == "nearest":        # utils.register_model(img_size, 'nearest') (ModeT/utils.py:74-83, infer.py:66)
            from .metrics import warp3d_nearest
            return warp3d_nearest(src, flow)
        if self.mode != "bilinear":
            raise NotImplementedError(f"SpatialTransformer: mode {self.mode!r} is not on the ModeT path")
        if torch.is_grad_enabled() and (src.requires_grad or flow.requires_grad):
            from . import autograd as ag
            return ag.Warp.apply(src, flow)
        return ops.warp3d(src, flow)


class ModeTransformer(nn.Module):
    """Reference ModeT/models.py:278-334 (and ModeT-cu/models.py:280-316, whose buffer is `v`)."""

    def __init__(self, dim, num_heads, kernel_size=3, qk_scale=None, use_rpb=True):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("ModeTransformer: the reference only ever uses kernel_size=3")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.kernel_size = kernel_size
        self.use_rpb = use_rpb
        if use_rpb:
            self.rpb = nn.Parameter(torch.zeros(num_heads, 3, 3, 3))
        off = torch.arange(-1, 2, dtype=torch.float32)
        self.register_buffer("grid", torch.stack(torch.meshgrid(off, off, off, indexing="ij"), -1))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # ModeT checkpoints carry the tap offsets as `grid` [3,3,3,3], ModeT-cu ones as `v` [27,3]: accept either, whichever
        # buffer this instance registers (ModeT: grid, ModeT_cu: v)
        mine, other = ("v", "grid") if "v" in self._buffers else ("grid", "v")
        t = state_dict.pop(prefix + other, None)
        if t is not None and prefix + mine not in state_dict:
            state_dict[prefix + mine] = t.reshape(27, 3) if mine == "v" else t.reshape(3, 3, 3, 3)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, q, k):
        return ops.modet_attention(q, k, self.rpb if self.use_rpb else None, self.num_heads, self.scale)


class ConvBlock(nn.Module):
    """Conv3d + LeakyReLU (reference ModeT/models.py:119-133)."""

    def __init__(self, in_channels, out_channels, kernal_size=3, stride=1, padding=1, alpha=0.1):
        super().__init__()
        _require_3x3(kernal_size, stride, padding, alpha)
        self.main = nn.Conv3d(in_channels, out_channels, kernal_size, stride, padding)
        self.activation = nn.LeakyReLU(alpha)

    def forward(self, x):
        return ops.conv3d(x, self.main.weight, self.main.bias, act_out=True)[0]


class ConvInsBlock(nn.Module):
    """Conv3d + InstanceNorm3d + LeakyReLU (reference ModeT/models.py:135-151)."""

    def __init__(self, in_channels, out_channels, kernal_size=3, stride=1, padding=1, alpha=0.1):
        super().__init__()
        _require_3x3(kernal_size, stride, padding, alpha)
        self.main = nn.Conv3d(in_channels, out_channels, kernal_size, stride, padding)
        self.norm = nn.InstanceNorm3d(out_channels)
        self.activation = nn.LeakyReLU(alpha)

    def raw(self, x, in_stats=None):
        """Raw conv output + its fp64 statistics; `in_stats` = statistics of a raw input."""
        return ops.conv3d(x, self.main.weight, self.main.bias, in_stats=in_stats, want_stats=True)

    def forward(self, x):
        raw, stats = self.raw(x)
        return ops.instnorm_lrelu_pool(raw, stats, pool=False, inplace=True)[0]


def _require_3x3(k, s, p, alpha):
    if (k, s, p) != (3, 1, 1) or abs(alpha - 0.1) > 1e-12:
        raise NotImplementedError("only kernel 3 / stride 1 / padding 1 / LeakyReLU(0.1) blocks are on the hot path")


class Encoder(nn.Module):
    """Reference ModeT/models.py:181-228: five-level conv pyramid, 2c,4c,8c,16c,32c channels."""

    def __init__(self, in_channel=1, first_out_channel=4):
        super().__init__()
        c = first_out_channel
        self.conv0 = nn.Sequential(ConvBlock(in_channel, c), ConvInsBlock(c, 2 * c), ConvInsBlock(2 * c, 2 * c))
        for lvl in range(1, 5):
            cin, cout = (2 ** lvl) * c, (2 ** (lvl + 1)) * c
            setattr(self, f"conv{lvl}", nn.Sequential(nn.AvgPool3d(2), ConvInsBlock(cin, cout), ConvInsBlock(cout, cout)))

    def forward_train(self, x):
        """Same graph through the autograd Functions (every forward and backward step is one of our kernels)."""
        from . import autograd as ag
        outs: List[torch.Tensor] = []
        c0 = self.conv0[0].main
        t = ag.ConvLReLU.apply(x, c0.weight, c0.bias)
        for lvl in range(5):
            seq = getattr(self, f"conv{lvl}")
            a = ag.ConvINLReLU.apply(t, seq[1].main.weight, seq[1].main.bias, False)
            if lvl < 4:
                out, t = ag.ConvINLReLU.apply(a, seq[2].main.weight, seq[2].main.bias, True)
            else:
                out = ag.ConvINLReLU.apply(a, seq[2].main.weight, seq[2].main.bias, False)
            outs.append(out)
        return tuple(outs)

    def forward(self, x):
        outs: List[torch.Tensor] = []
        t = self.conv0[0](x)
        for lvl in range(5):
            seq = getattr(self, f"conv{lvl}")
            raw, st = seq[1].raw(t)                       # conv (+ stats)
            raw, st = seq[2].raw(raw, in_stats=st)        # IN+LReLU on load, conv (+ stats)
            out, t = ops.instnorm_lrelu_pool(raw, st, pool=lvl < 4, inplace=True)
            outs.append(out)
        return tuple(outs)


class ProjectionLayer(nn.Module):
    """Reference ModeT/models.py:230-241."""

    def __init__(self, in_channels, dim=6, norm=nn.LayerNorm):
        super().__init__()
        self.norm = norm(dim)
        self.proj = nn.Linear(in_channels, dim)
        self.proj.weight = nn.Parameter(Normal(0, 1e-5).sample(self.proj.weight.shape))
        self.proj.bias = nn.Parameter(torch.zeros(self.proj.bias.shape))

    def forward(self, feat):
        if not isinstance(self.norm, nn.LayerNorm):
            raise NotImplementedError("ProjectionLayer: only norm=nn.LayerNorm is on the hot path")
        return ops.proj_ln(feat, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias, self.norm.eps)

    def of_warped(self, feat, flow):
        """self(SpatialTransformer(feat, flow)) without materialising the warped volume (models.py:388-389)."""
        return ops.warp_proj_ln(feat, flow, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias,
                                self.norm.eps)


class CWM(nn.Module):
    """Competitive weighting module, reference ModeT/models.py:243-275."""

    def __init__(self, in_channels, channels):
        super().__init__()
        self.num_fields = in_channels // 3
        self.conv = nn.Sequential(ConvInsBlock(in_channels, channels, 3, 1), ConvInsBlock(channels, channels, 3, 1),
                                  nn.Conv3d(channels, self.num_fields, 3, 1, 1), nn.Softmax(dim=1))
        self.upsample = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)

    def forward_train(self, x):
        from . import autograd as ag
        u = ag.Upsample2x.apply(x, 1.0)
        a = ag.ConvINLReLU.apply(u, self.conv[0].main.weight, self.conv[0].main.bias, False)
        a = ag.ConvINLReLU.apply(a, self.conv[1].main.weight, self.conv[1].main.bias, False)
        logits = ag.Conv.apply(a, self.conv[2].weight, self.conv[2].bias)
        return ag.CwmFuse.apply(u, logits)

    def forward(self, x):
        u = ops.upsample2x(x)
        raw, st = self.conv[0].raw(u)
        raw, st = self.conv[1].raw(raw, in_stats=st)
        logits, _ = ops.conv3d(raw, self.conv[2].weight, self.conv[2].bias, in_stats=st)
        return ops.cwm_fuse(u, logits)


class ModeT(nn.Module):
    """Reference ModeT/models.py:337-412: encoder + five-level motion-decomposition decoder."""

    def __init__(self, inshape=(160, 192, 160), in_channel=1, channels=4, head_dim=6, num_heads=[8, 4, 2, 1, 1],
                 scale=None):
        super().__init__()
        if num_heads[3] != 1 or num_heads[4] != 1:
            raise ValueError("levels 2 and 1 have no CWM in ModeT, so num_heads[3] and num_heads[4] must be 1")
        self.channels = channels
        self.step = 7
        self.inshape = inshape
        # "fp32" (reference numerics) or "bf16": Conv3d products on bf16 tensor cores with fp32 accumulation
        # (BASELINE.json configs[2..3]); not part of the reference constructor, set as an attribute
        self.conv_precision = "fp32"
        c = channels
        self.encoder = Encoder(in_channel=in_channel, first_out_channel=c)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.upsample_trilin = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)
        for level in range(1, 6):
            heads = num_heads[5 - level]
            setattr(self, f"projblock{level}", ProjectionLayer((2 ** level) * c, dim=head_dim * heads))
            setattr(self, f"mdt{level}", ModeTransformer(head_dim * heads, heads, qk_scale=scale))
            if level >= 3:
                setattr(self, f"cwm{level}", CWM(3 * heads, 3 * heads * 2))
        self.transformer = nn.ModuleList(SpatialTransformer([s // 2 ** i for s in inshape]) for i in range(4))

    def _attend(self, level: int, feat_f, feat_m, flow):
        """mdt(proj(F), proj(T(M, flow))) -- models.py:388-390."""
        pb, mdt = getattr(self, f"projblock{level}"), getattr(self, f"mdt{level}")
        return mdt(pb(feat_f), pb.of_warped(feat_m, flow))

    def _forward_train(self, moving, fixed):
        """Training graph (models.py:377-412) through smilecode_b200.autograd: unfused ops so that every node has a
        hand-written backward kernel; the tensor glue (cat / slicing / add / scalar multiply) is torch's."""
        from . import autograd as ag
        B = moving.shape[0]
        feats = self.encoder.forward_train(torch.cat([moving, fixed], 0))
        M = [f[:B] for f in feats]
        Fx = [f[B:] for f in feats]

        def proj(level, feat):
            pb = getattr(self, f"projblock{level}")
            return ag.ProjLN.apply(feat, pb.proj.weight, pb.proj.bias, pb.norm.weight, pb.norm.bias, pb.norm.eps)

        def attend(level, feat_f, feat_m):
            mdt = getattr(self, f"mdt{level}")
            return ag.Attention.apply(proj(level, feat_f), proj(level, feat_m), mdt.rpb if mdt.use_rpb else None,
                                      mdt.num_heads, mdt.scale)

        flow = self.cwm5.forward_train(attend(5, Fx[4], M[4]))
        w = self.cwm4.forward_train(attend(4, Fx[3], ag.Warp.apply(M[3], flow)))
        flow = ag.Warp.apply(ag.Upsample2x.apply(flow, 2.0), w) + w
        w = self.cwm3.forward_train(attend(3, Fx[2], ag.Warp.apply(M[2], flow)))
        flow = ag.Warp.apply(ag.Upsample2x.apply(flow, 2.0), w) + w
        w = attend(2, Fx[1], ag.Warp.apply(M[1], flow))
        flow = ag.Upsample2x.apply(ag.Warp.apply(flow, w) + w, 2.0)
        w = attend(1, Fx[0], ag.Warp.apply(M[0], flow))
        flow = ag.Warp.apply(flow, w) + w
        return ag.Warp.apply(moving, flow), flow

    # The tensor-core convolutions keep prepared copies of their weights (ops._prepared_weights); every entry point
    # through which weights are commonly replaced drops them.
    def load_state_dict(self, *args, **kwargs):
        ops.invalidate_prepared_weights()
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode: bool = True):
        ops.invalidate_prepared_weights()
        return super().train(mode)

    def forward(self, moving, fixed):
        with ops.conv_precision(self.conv_precision):
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                return self._forward_train(moving, fixed)
            with ops.stats_arena(moving.device):
                return self._forward_infer(moving, fixed)

    def _forward_infer(self, moving, fixed):
        B = moving.shape[0]
        feats = self.encoder(torch.cat([moving, fixed], 0))     # shared weights: one batched pass
        M = [f[:B] for f in feats]
        Fx = [f[B:] for f in feats]

        # level 5 .. 3: multi-head attention -> CWM fusion (models.py:383-398)
        qk = self.projblock5(feats[4])                           # both volumes in one launch
        flow = self.cwm5(self.mdt5(qk[B:], qk[:B]))
        w = self.cwm4(self._attend(4, Fx[3], M[3], flow))
        flow = ops.flow_compose(ops.upsample2x(flow, 2.0), w)
        w = self.cwm3(self._attend(3, Fx[2], M[2], flow))
        flow = ops.flow_compose(ops.upsample2x(flow, 2.0), w)

        if not FUSE_LEVELS:
            w = self._attend(2, Fx[1], M[1], flow)
            flow = ops.upsample2x(ops.flow_compose(flow, w), 2.0)
            w = self._attend(1, Fx[0], M[0], flow)
            flow = ops.flow_compose(flow, w)
            return ops.warp3d(moving, flow), flow
        # level 2, 1: single head, attention + compose (+ final warp) fused (models.py:400-410)
        pb, mdt = self.projblock2, self.mdt2
        f2, _ = ops.modet_fused(pb(Fx[1]), pb.of_warped(M[1], flow), mdt.rpb if mdt.use_rpb else None, flow, None,
                                mdt.scale, postmul=2.0, ln_gamma=pb.norm.weight, ln_beta=pb.norm.bias)
        flow = ops.upsample2x(f2)
        pb, mdt = self.projblock1, self.mdt1
        flow, y_moved = ops.modet_fused(pb(Fx[0]), pb.of_warped(M[0], flow), mdt.rpb if mdt.use_rpb else None, flow,
                                        moving, mdt.scale, postmul=1.0, ln_gamma=pb.norm.weight, ln_beta=pb.norm.bias)
        return y_moved, flow


class ModeT_cu(ModeT):
    """Name used by the reference's ModeT-cu/train.py:14 (default scale=1, ModeT-cu/models.py:325).  Its state_dict carries
    the tap offsets as `mdtN.v` [27,3] like ModeT-cu/models.py:296-299 (not `mdtN.grid` [3,3,3,3] as ModeT does), so a
    checkpoint written here loads with strict=True into the reference's ModeT_cu and vice versa; `grid` is accepted on
    load too (ModeTransformer._load_from_state_dict handles both spellings)."""

    def __init__(self, inshape=(160, 192, 160), in_channel=1, channels=4, head_dim=6, num_heads=[8, 4, 2, 1, 1], scale=1):
        super().__init__(inshape, in_channel, channels, head_dim, num_heads, scale)
        for level in range(1, 6):
            mdt = getattr(self, f"mdt{level}")
            g = mdt.grid
            del mdt._buffers["grid"]
            mdt.register_buffer("v", g.reshape(27, 3).clone())


# ------------------------------------------------------------------------------------------------
# Classes the reference module also exports but never calls from ModeT.forward (SURVEY.md section 2,
# row 16).  Kept importable with the same constructor signatures; they are off the hot path.
# ------------------------------------------------------------------------------------------------
class VecInt(nn.Module):
    """Scaling-and-squaring integration (reference ModeT/models.py:70-87); built on our warp."""

    def __init__(self, inshape, nsteps=7):
        super().__init__()
        assert nsteps >= 0, "nsteps should be >= 0, found: %d" % nsteps
        self.nsteps = nsteps
        self.scale = 1.0 / (2 ** nsteps)
        self.transformer = SpatialTransformer(inshape)

    def forward(self, vec):
        vec = vec * self.scale
        for _ in range(self.nsteps):
            vec = vec + self.transformer(vec, vec)
        return vec


class ResizeTransform(nn.Module):
    """Resize + rescale a vector field (reference ModeT/models.py:90-116)."""

    def __init__(self, vel_resize, ndims):
        super().__init__()
        self.factor = 1.0 / vel_resize
        self.mode = {2: "bilinear", 3: "trilinear"}.get(ndims, "linear")

    def forward(self, x):
        if self.factor == 1:
            return x
        y = nnf.interpolate(x, align_corners=True, scale_factor=self.factor, mode=self.mode)
        return self.factor * y


class UpConvBlock(nn.Module):
    """Reference ModeT/models.py:153-166 (unused by ModeT.forward)."""

    def __init__(self, in_channels, out_channels, kernel_size=4, stride=2, alpha=0.1):
        super().__init__()
        self.upconv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=kernel_size, stride=stride)
        self.actout = nn.Sequential(nn.InstanceNorm3d(out_channels), nn.LeakyReLU(alpha))

    def forward(self, x):
        return self.actout(self.upconv(x)[:, :, 1:-1, 1:-1, 1:-1])


class DeconvBlock(nn.Module):
    """Reference ModeT/models.py:168-179 (unused by ModeT.forward)."""

    def __init__(self, dec_channels, skip_channels):
        super().__init__()
        self.upconv = UpConvBlock(dec_channels, skip_channels)
        self.conv = nn.Sequential(ConvInsBlock(2 * skip_channels, skip_channels), ConvInsBlock(skip_channels, skip_channels))

    def forward(self, dec, skip):
        return self.conv(torch.cat([self.upconv(dec), skip], dim=1))
