#!/usr/bin/env python
"""Which stage of the CUDA forward costs the distance to the exact (fp64) answer at full size?  

    python tools/parity_bisect.py [DxHxW] [h5,h4,h3,h2,h1]

Computes the reference path (oracle port, checker only) in fp32 and fp64 once, then runs the CUDA forward in child
processes under the bisect knobs (they are read once per process):
    default                 the shipped configuration
    conv_simt               SMILE_CONV_TC=0        : all convolutions on the fp32 SIMT kernels (no 3xTF32 tcgen05)
    unfused                 models.FUSE_LEVELS=0   : L2/L1 as attention -> flow_compose -> warp3d (bit-exact warp kernels)
    unfused_exact           + SMILE_ATTN_EXACT=1   : generic attention with exp2f and a true division (no ex2.approx)
    unfused_exact_simt      + SMILE_CONV_TC=0      : everything fp32-exact-class
and prints, for each, max / RMS / p99.9 of |flow - ref64| next to the reference's own |ref32 - ref64|.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "default": {},
    "conv_simt": {"SMILE_CONV_TC": "0"},
    "unfused": {"SMILE_BISECT_UNFUSED": "1"},
    "unfused_exact": {"SMILE_BISECT_UNFUSED": "1", "SMILE_ATTN_EXACT": "1"},
    "unfused_exact_simt": {"SMILE_BISECT_UNFUSED": "1", "SMILE_ATTN_EXACT": "1", "SMILE_CONV_TC": "0"},
}


def stats(a, ref64):
    import torch
    d = (a.double() - ref64).abs().flatten()
    sub = d[::5]
    i = int(d.argmax())
    return {"max": float(d.max()), "argmax": i, "rms": float(d.pow(2).mean().sqrt()),
            "p999": float(sub.kthvalue(int(0.999 * sub.numel())).values)}


def child(shape, heads, cache):
    import torch
    from oracle import modet_oracle as orc          # checker only
    from smilecode_b200 import models
    from smilecode_b200.synth import make_pair
    if os.environ.get("SMILE_BISECT_UNFUSED") == "1":
        models.FUSE_LEVELS = False
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    moving, fixed = make_pair(shape, batch=1, seed=24)
    model = models.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    with torch.no_grad():
        _, flow = model(moving.cuda(), fixed.cuda())
    flow = flow.cpu()
    ref = torch.load(cache)
    s = stats(flow, ref["f64"])
    s["max_vs_ref32"] = float((flow - ref["f32"]).abs().max())
    s["ref32_err_at_argmax"] = float((ref["f32"].double() - ref["f64"]).abs().flatten()[s["argmax"]])
    print("BISECT " + json.dumps(s))


def main():
    shape = tuple(int(x) for x in sys.argv[1].split("x")) if len(sys.argv) > 1 else (160, 192, 160)
    heads = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 4, 2, 1, 1]
    cache = f"/tmp/smile_bisect_{'x'.join(map(str, shape))}_{'_'.join(map(str, heads))}.pt"
    if os.environ.get("SMILE_BISECT_CHILD"):
        return child(shape, heads, cache)
    import torch
    from oracle import modet_oracle as orc
    from smilecode_b200.synth import make_pair
    torch.set_num_threads(os.cpu_count() or 1)
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    moving, fixed = make_pair(shape, batch=1, seed=24)
    with torch.no_grad():
        _, f32 = orc.modet_forward(moving, fixed, sd, num_heads=heads, scale=1.0, library_ops=True)
        sd64 = {k: v.double() for k, v in sd.items()}
        _, f64 = orc.modet_forward(moving.double(), fixed.double(), sd64, num_heads=heads, scale=1.0, library_ops=True)
    torch.save({"f32": f32, "f64": f64}, cache)
    r = stats(f32, f64)
    print(f"shape {shape} heads {heads}  |flow|max {float(f64.abs().max()):.3f}")
    print(f"{'reference fp32':22s} max {r['max']:.3e}  rms {r['rms']:.3e}  p99.9 {r['p999']:.3e}   (vs fp64; the noise floor)")
    for name, env in VARIANTS.items():
        e = dict(os.environ, SMILE_BISECT_CHILD="1", **env)
        out = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=e, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("BISECT ")]
        if not line:
            print(f"{name:22s} FAILED\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
            continue
        s = json.loads(line[0][7:])
        print(f"{name:22s} max {s['max']:.3e} ({s['max'] / r['max']:.2f}x)  rms {s['rms']:.3e} ({s['rms'] / r['rms']:.2f}x)  "
              f"p99.9 {s['p999']:.3e} ({s['p999'] / r['p999']:.2f}x)   |ours-ref32|max {s['max_vs_ref32']:.3e}   "
              f"ref32's own error at our worst voxel {s['ref32_err_at_argmax']:.3e}")
    os.remove(cache)


if __name__ == "__main__":
    main()
