"""Pins oracle/modet_oracle.py against fixtures produced by the reference's own modules
(oracle/make_golden.py).  CPU only."""
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import modet_oracle as orc
from smilecode_b200.synth import make_pair


def _params(g):
    return {k[2:]: v for k, v in g.items() if k.startswith("p.")}


@pytest.mark.parametrize("name", golden_names("attn_"))
def test_attention(name):
    g = load_golden(name)
    out = orc.modet_attention(g["q"], g["k"], g["rpb"], int(g["heads"]), float(g["scale"]))
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max() <= 2e-6


@pytest.mark.parametrize("name", golden_names("warp_"))
def test_warp_bit_exact(name):
    g = load_golden(name)
    out = orc.warp_trilinear(g["src"], g["flow"])
    # identical bits to torch grid_sample => floor() corner indices identical too
    assert torch.equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names("up2_"))
def test_upsample(name):
    g = load_golden(name)
    out = orc.upsample2x_trilinear(g["x"])
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max() <= 1e-6


@pytest.mark.parametrize("name", golden_names("cwm_"))
def test_cwm(name):
    g = load_golden(name)
    sd = {"cwm." + k: v for k, v in _params(g).items()}
    out = orc.cwm(g["x"], sd, "cwm")
    assert (out - g["out"]).abs().max() <= 5e-6


@pytest.mark.parametrize("name", golden_names("proj_"))
def test_projection(name):
    g = load_golden(name)
    sd = {"pb." + k: v for k, v in _params(g).items()}
    out = orc.projection(g["x"], sd, "pb")
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max() <= 5e-6


def test_encoder():
    g = load_golden("encoder_b2_16x16x32")
    sd = orc.synth_state_dict(seed=1234)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - g["weights_checksum"]) < 1e-6
    outs = orc.encoder(g["x"], sd)
    for i, o in enumerate(outs):
        # the last level is 1x1x2 voxels: InstanceNorm over two samples is ill-conditioned
        assert (o - g[f"out{i}"]).abs().max() <= (2e-5 if i < 4 else 1e-4), i


@pytest.mark.parametrize("name", golden_names("e2e_"))
def test_end_to_end(name):
    g = load_golden(name)
    heads = [int(h) for h in g["num_heads"]]
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - g["weights_checksum"]) < 1e-6
    shape = tuple(g["flow"].shape[2:])
    moving, fixed = make_pair(shape, batch=1, seed=24)
    assert abs(float(moving.double().sum()) - g["moving_checksum"]) < 1e-6
    moved, flow = orc.modet_forward(moving, fixed, sd, num_heads=heads, scale=1.0)
    # the reference's own fp32-vs-fp64 noise floor on this draw is ~2e-5 (stored as flow_fp64)
    floor = (g["flow"] - g["flow_fp64"]).abs().max()
    assert floor < 1e-4
    assert (flow - g["flow"]).abs().max() <= 1e-4
    assert (moved - g["moved"]).abs().max() <= 1e-4


def test_losses():
    g = load_golden("losses_12x14x11")
    assert abs(float(orc.ncc_vxm(g["a"], g["b"])) - g["ncc"]) <= 1e-6
    assert abs(float(orc.grad3d_l2(g["flow"])) - g["grad"]) <= 1e-7


def test_library_ops_variant_matches():
    """The timing variant of the oracle (torch grid_sample / interpolate, as the reference calls them)
    must agree with the elementary restatement."""
    heads = [8, 4, 2, 1, 1]
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    moving, fixed = make_pair((32, 32, 32), batch=1, seed=24)
    a = orc.modet_forward(moving, fixed, sd, num_heads=heads, scale=1.0)
    b = orc.modet_forward(moving, fixed, sd, num_heads=heads, scale=1.0, library_ops=True)
    assert (a[1] - b[1]).abs().max() <= 2e-5 and (a[0] - b[0]).abs().max() <= 2e-5
