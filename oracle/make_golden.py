"""Generate tests/golden/*.npz by running the REFERENCE's own modules (ModeT/models.py,
ModeT/losses.py, imported unmodified from /root/reference) on seeded inputs, CPU fp32.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The fixtures pin oracle/modet_oracle.py (tests/test_oracle_golden.py) and, through it and
directly, the CUDA path (tests/test_gpu_parity.py).  Test infrastructure, not product code.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("SMILE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "ModeT"))
warnings.filterwarnings("ignore")

import models as ref_models  # noqa: E402  (the reference)
import losses as ref_losses  # noqa: E402

from oracle import modet_oracle as orc  # noqa: E402
from smilecode_b200.synth import make_pair  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                           for k, v in arrs.items()})
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrs.items()})


def smooth_flow(g, B, shape, amp):
    cs = [max(2, s // 4) for s in shape]
    c = (torch.rand(B, 3, *cs, generator=g) * 2 - 1) * amp
    return torch.nn.functional.interpolate(c, size=tuple(shape), mode="trilinear", align_corners=True).contiguous()


@torch.no_grad()
def main():
    g = torch.Generator().manual_seed(2024)

    # ---- a2 ModeTransformer -------------------------------------------------------------
    for tag, (B, D, H, W, heads, hd, scale) in {
        "attn_b2_5x6x7_h4": (2, 5, 6, 7, 4, 6, None),
        "attn_b1_2x2x2_h8": (1, 2, 2, 2, 8, 6, 1.0),
        "attn_b1_8x6x12_h1": (1, 8, 6, 12, 1, 6, 1.0),
        "attn_b1_3x1x4_h2_hd4": (1, 3, 1, 4, 2, 4, 0.7),
    }.items():
        m = ref_models.ModeTransformer(heads * hd, heads, qk_scale=scale)
        m.rpb.data = torch.randn(m.rpb.shape, generator=g) * 0.5
        q = torch.randn(B, D, H, W, heads * hd, generator=g)
        k = torch.randn(B, D, H, W, heads * hd, generator=g)
        save(tag, q=q, k=k, rpb=m.rpb.data, heads=heads, scale=float(m.scale), out=m(q, k))

    # ---- a5 SpatialTransformer ------------------------------------------------------------
    for tag, (B, C, shape, amp) in {
        "warp_b2_c3_5x6x7": (2, 3, (5, 6, 7), 2.5),
        "warp_b1_c5_10x12x7": (1, 5, (10, 12, 7), 4.0),
        "warp_b1_c2_20x24x20_zero": (1, 2, (20, 24, 20), 0.0),
        "warp_b1_c1_14x9x11_big": (1, 1, (14, 9, 11), 20.0),
    }.items():
        st = ref_models.SpatialTransformer(shape)
        src = torch.randn(B, C, *shape, generator=g)
        flow = smooth_flow(g, B, shape, amp) + 0.3 * torch.randn(B, 3, *shape, generator=g) * (amp > 0)
        save(tag, src=src, flow=flow, out=st(src, flow))

    # ---- a6 trilinear x2 upsample ---------------------------------------------------------
    up = torch.nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)
    for tag, (B, C, shape) in {"up2_b2_c3_5x6x7": (2, 3, (5, 6, 7)), "up2_b1_c24_2x2x2": (1, 24, (2, 2, 2)),
                               "up2_b1_c3_10x12x10": (1, 3, (10, 12, 10))}.items():
        x = torch.randn(B, C, *shape, generator=g)
        save(tag, x=x, out=up(x))

    # ---- a4 CWM ---------------------------------------------------------------------------
    for tag, (B, heads, shape) in {"cwm_b2_h2_5x6x7": (2, 2, (5, 6, 7)), "cwm_b1_h8_2x3x2": (1, 8, (2, 3, 2))}.items():
        torch.manual_seed(7)
        m = ref_models.CWM(3 * heads, 6 * heads)
        x = torch.rand(B, 3 * heads, *shape, generator=g) * 2 - 1
        sd = {k: v for k, v in m.state_dict().items()}
        save(tag, x=x, out=m(x), **{"p." + k: v for k, v in sd.items()})

    # ---- a7 ProjectionLayer ---------------------------------------------------------------
    for tag, (B, cin, dim, shape) in {"proj_b2_8to6_5x6x7": (2, 8, 6, (5, 6, 7)),
                                      "proj_b1_128to48_2x3x2": (1, 128, 48, (2, 3, 2))}.items():
        m = ref_models.ProjectionLayer(cin, dim=dim)
        m.proj.weight.data = torch.randn(dim, cin, generator=g) / cin ** 0.5
        m.proj.bias.data = torch.randn(dim, generator=g) * 0.1
        m.norm.weight.data = torch.rand(dim, generator=g) + 0.5
        m.norm.bias.data = torch.randn(dim, generator=g) * 0.1
        x = torch.randn(B, cin, *shape, generator=g)
        save(tag, x=x, out=m(x), **{"p." + k: v for k, v in m.state_dict().items()})

    # ---- a8 Encoder -----------------------------------------------------------------------
    enc = ref_models.Encoder(in_channel=1, first_out_channel=4)
    sd = orc.synth_state_dict(seed=1234)     # weights are rebuilt from the seed, not stored
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    x = torch.rand(2, 1, 16, 16, 32, generator=g)
    outs = enc(x)
    save("encoder_b2_16x16x32", x=x, **{f"out{i}": o for i, o in enumerate(outs)},
         weights_checksum=float(sum(v.double().sum() for v in sd.values())))

    # ---- a9 ModeT end to end (infer.py:61-86 plumbing with synthetic tensors) ----------------
    for tag, shape, heads in (("e2e_32x32x32", (32, 32, 32), [8, 4, 2, 1, 1]),
                              ("e2e_32x48x32_h6", (32, 48, 32), [6, 3, 2, 1, 1])):
        sd = orc.synth_state_dict(seed=1234, num_heads=heads)
        model = ref_models.ModeT(shape, head_dim=6, num_heads=heads, scale=1).eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("grid") for k in missing), (missing, unexpected)
        moving, fixed = make_pair(shape, batch=1, seed=24)
        moved, flow = model(moving, fixed)
        m64 = ref_models.ModeT(shape, head_dim=6, num_heads=heads, scale=1).eval()
        m64.load_state_dict(sd, strict=False)
        m64 = m64.double()
        moved64, flow64 = m64(moving.double(), fixed.double())
        csum = float(sum(v.double().sum() for v in sd.values()))
        save(tag, moved=moved, flow=flow, flow_fp64=flow64.float(),
             weights_checksum=csum, num_heads=np.asarray(heads),
             moving_checksum=float(moving.double().sum()), fixed_checksum=float(fixed.double().sum()))
        print(tag, "|flow|max", float(flow.abs().max()), "ref fp32-vs-fp64 flow max-abs",
              float((flow.double() - flow64).abs().max()))

    # ---- a10 losses (NCC_vxm hard-codes .to('cuda') at losses.py:57 -> neutralised here) ---
    class _Stay(torch.Tensor):
        def to(self, *a, **k):
            return torch.Tensor(self)

    real_ones = torch.ones
    ref_losses.torch.ones = lambda *a, **k: real_ones(*a, **k).as_subclass(_Stay)
    try:
        a = torch.rand(2, 1, 12, 14, 11, generator=g)
        b = (a + 0.2 * torch.rand(2, 1, 12, 14, 11, generator=g)).clamp(0, 1)
        ncc = ref_losses.NCC_vxm()(a, b)
    finally:
        ref_losses.torch.ones = real_ones
    fl = smooth_flow(g, 2, (12, 14, 11), 2.0)
    gr = ref_losses.Grad3d(penalty="l2")(fl, None)
    save("losses_12x14x11", a=a, b=b, ncc=ncc, flow=fl, grad=gr)


if __name__ == "__main__":
    main()
