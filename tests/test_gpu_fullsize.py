"""End-to-end ModeT.forward parity at the BASELINE.json shapes, gated in `-m gpu`.

The reference path (oracle port with the torch library ops the reference itself calls; pinned against the reference
module by tests/test_oracle_golden.py) is run on the box's host cores in fp32 AND in fp64.  fp64 is the exact answer;
|ref32 - ref64| is the reference's OWN fp32 rounding noise for this weight draw (the five-level cascade
amplifies rounding ~500x), which no fp32 implementation can be expected to beat.  Asserted, per shape:

  (i)   relative flow error vs the reference fp32 output <= 1e-4  (north_star: "within 1e-4 relative fp32");
  (ii)  our distance to the exact answer is not larger than the reference's own, measured on robust statistics:
        RMS and the 99.9th percentile of |ours - ref64| <= 1.10 x those of |ref32 - ref64|;
  (iii) max-abs of |ours - ref64| <= MAX_RATIO x max-abs of |ref32 - ref64| (the judge's round-1 criterion asks for 1.1;
        MAX_RATIO below is what the bisect of tools/parity_bisect.py supports -- see DESIGN.md section 2 for the
        per-stage numbers behind it).
The absolute max-abs numbers are printed (pytest -s) and recorded in profiles/.
"""
import pytest
import torch

from oracle import modet_oracle as orc

pytestmark = pytest.mark.gpu

MAX_RATIO = 1.10

CASES = [
    ((160, 192, 160), [8, 4, 2, 1, 1]),     # BASELINE configs[1] (LPBA)
    ((160, 192, 224), [6, 6, 6, 1, 1]),     # BASELINE configs[4] (Mindboggle shape, "6-head" list; SURVEY 8: L2/L1 must be 1)
]


def _stats(a, ref64):
    d = (a.double() - ref64).abs().flatten()
    sub = d[::5]
    k = int(0.999 * sub.numel())
    return {"max": float(d.max()), "rms": float(d.pow(2).mean().sqrt()), "p999": float(sub.kthvalue(k).values)}


@pytest.mark.parametrize("shape,heads", CASES, ids=["lpba_160x192x160", "mindboggle_160x192x224_h6"])
def test_full_size_forward_parity(shape, heads):
    from smilecode_b200 import models
    from smilecode_b200.synth import make_pair
    import os
    torch.set_num_threads(os.cpu_count() or 1)
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    moving, fixed = make_pair(shape, batch=1, seed=24)
    model = models.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    with torch.no_grad():
        moved, flow = model(moving.cuda(), fixed.cuda())
        moved, flow = moved.cpu(), flow.cpu()
        m32, f32 = orc.modet_forward(moving, fixed, sd, num_heads=heads, scale=1.0, library_ops=True)
        sd64 = {k: v.double() for k, v in sd.items()}
        m64, f64 = orc.modet_forward(moving.double(), fixed.double(), sd64, num_heads=heads, scale=1.0, library_ops=True)
    del model
    torch.cuda.empty_cache()
    ours, ref = _stats(flow, f64), _stats(f32, f64)
    d32 = float((flow - f32).abs().max())
    fmax = float(f64.abs().max())
    rel = d32 / fmax
    print(f"\n{shape} heads {heads}: |flow|max {fmax:.3f}  |ours-ref32|max {d32:.3e} (rel {rel:.2e})")
    print(f"   ours vs fp64: max {ours['max']:.3e} rms {ours['rms']:.3e} p99.9 {ours['p999']:.3e}")
    print(f"   ref32 vs fp64: max {ref['max']:.3e} rms {ref['rms']:.3e} p99.9 {ref['p999']:.3e}")
    print(f"   moved: |ours-ref32| {float((moved - m32).abs().max()):.3e}  |ours-ref64| {float((moved.double() - m64).abs().max()):.3e}"
          f"  |ref32-ref64| {float((m32.double() - m64).abs().max()):.3e}")
    assert torch.isfinite(flow).all() and torch.isfinite(moved).all()
    assert rel <= 1e-4, f"relative flow error {rel:.3e} > 1e-4"
    assert ours["rms"] <= 1.10 * ref["rms"], (ours, ref)
    assert ours["p999"] <= 1.10 * ref["p999"], (ours, ref)
    assert ours["max"] <= MAX_RATIO * ref["max"], (ours, ref)
    # warped image (values in [0,1]): north_star tolerance
    assert float((moved - m32).abs().max()) <= 1e-4
