#!/usr/bin/env python
"""GPU comparators on the B200 (SURVEY 8d last paragraph) -- informational.

    python tools/comparators.py [--shape DxHxW] [--json out.json]

Times, with CUDA events on the current stream, after warm-up, on the same synthetic pair and weights:
  * ours            smilecode_b200.models.ModeT                           (forward, fp32)
  * ref_eager       the reference's ModeT/models.py in PyTorch eager on the GPU, TF32 off
  * ref_cu          the reference's ModeT-cu/models.py with its own `modet` extension built for sm_100
  * a3 micro-bench  smile_modet_qkrpb_fwd / _bwd  vs  the reference extension's modet_fw / modet_bw at L1
The reference tree comes from baseline/_ref (staged by oracle/stage_reference.py); rows whose pieces are missing are
reported as null.  Nothing here is on the product path.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def time_ms(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="160x192x160")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    shape = tuple(int(x) for x in args.shape.split("x"))
    heads = [8, 4, 2, 1, 1]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    from oracle import reference_loader as rl       # checker-side helper
    from oracle import modet_oracle as orc
    from smilecode_b200 import models, ops
    from smilecode_b200.synth import make_pair
    dev = torch.device("cuda", 0)
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    moving, fixed = (t.to(dev) for t in make_pair(shape, batch=1, seed=24))
    res = {"shape": list(shape), "heads": heads, "gpu": torch.cuda.get_device_name(0)}

    ours = models.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
    ours.load_state_dict(sd, strict=False)
    ours = ours.to(dev).eval()
    with torch.no_grad():
        res["ours_ms"] = time_ms(lambda: ours(moving, fixed), warm=3, iters=10)
        y_o, f_o = ours(moving, fixed)

    ref = rl.reference_models()
    res["ref_eager_ms"] = res["ref_cu_ms"] = None
    if ref is not None:
        m = ref.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
        m.load_state_dict(sd, strict=False)
        m = m.to(dev).eval()
        with torch.no_grad():
            res["ref_eager_ms"] = time_ms(lambda: m(moving, fixed), warm=2, iters=3)
            y_r, f_r = m(moving, fixed)
        res["ref_eager_vs_ours_max_abs_flow"] = float((f_r - f_o).abs().max())
        res["ref_eager_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
        del m, y_r, f_r
        torch.cuda.empty_cache()
    refcu = rl.reference_models_cu()
    if refcu is not None:
        m = refcu.ModeT_cu(shape, head_dim=6, num_heads=heads, scale=1)
        sd_cu = {k: v for k, v in sd.items()}
        m.load_state_dict(sd_cu, strict=False)
        m = m.to(dev).eval()
        with torch.no_grad():
            res["ref_cu_ms"] = time_ms(lambda: m(moving, fixed), warm=2, iters=3)
            y_r, f_r = m(moving, fixed)
        res["ref_cu_vs_ours_max_abs_flow"] = float((f_r - f_o).abs().max())
        del m, y_r, f_r
        torch.cuda.empty_cache()

        # ---- a3 micro-benchmark at L1: same tensors, same layouts (modet.cpp:4-37)
        import importlib
        sys.path[:0] = [os.path.join(rl.staged_dir(), "ModeT-cu", "modet")]
        modet = importlib.import_module("modet")
        D, H, W = shape
        g = torch.Generator(device=dev).manual_seed(3)
        q = torch.randn(1, 1, D, H, W, 6, device=dev, generator=g)
        k = torch.nn.functional.pad(torch.randn(1, 1, D, H, W, 6, device=dev, generator=g), (0, 0, 1, 1, 1, 1, 1, 1)).contiguous()
        rpb = torch.randn(1, 3, 3, 3, device=dev, generator=g)
        d_attn = torch.randn(1, 1, D, H, W, 27, device=dev, generator=g)
        a_ref = modet.modet_fw(q, k, rpb)
        a_our = ops.modet_qkrpb_fwd(q, k, rpb)
        res["qkrpb_fwd_max_abs_diff"] = float((a_ref - a_our).abs().max())
        res["qkrpb_fwd_ms"] = {"reference_modet_fw": time_ms(lambda: modet.modet_fw(q, k, rpb)),
                               "ours": time_ms(lambda: ops.modet_qkrpb_fwd(q, k, rpb))}
        gr = modet.modet_bw(d_attn, q, k, True)
        go = ops.modet_qkrpb_bwd(d_attn, q, k, True)
        res["qkrpb_bwd_max_abs_diff"] = [float((a - b.reshape(a.shape)).abs().max()) for a, b in zip(gr, go)]
        res["qkrpb_bwd_ms"] = {"reference_modet_bw": time_ms(lambda: modet.modet_bw(d_attn, q, k, True)),
                               "ours": time_ms(lambda: ops.modet_qkrpb_bwd(d_attn, q, k, True))}
        # algorithmic bytes: fwd reads q 24 + k 24 (+halo), writes 108 B/voxel; bwd reads 108+48, writes 48
        N = D * H * W
        res["qkrpb_fwd_gbs_ours"] = 156 * N / res["qkrpb_fwd_ms"]["ours"] / 1e6
        res["qkrpb_bwd_gbs_ours"] = 204 * N / res["qkrpb_bwd_ms"]["ours"] / 1e6
    print(json.dumps(res, indent=1))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
