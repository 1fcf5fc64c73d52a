"""Evaluation helpers of ModeT/infer.py: oracle vs reference-generated goldens (CPU), CUDA kernels vs
the oracle and the goldens through the C ABI (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as morc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["metrics_a", "metrics_b"]


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_goldens(name):
    g = load(name)
    flow, seg_m, seg_f = torch.from_numpy(g["flow"]), torch.from_numpy(g["seg_m"]), torch.from_numpy(g["seg_f"])
    warped = morc.warp_nearest(seg_m, flow)
    assert np.array_equal(warped.numpy(), g["warped"])
    assert morc.dice_val_VOI(warped.long().numpy()[0, 0], seg_f.long().numpy()[0, 0]) == float(g["dsc"])
    det = morc.jacobian_determinant_vxm(g["flow"][0])
    assert np.array_equal(det, g["det"])
    assert int(np.sum(det <= 0)) == int(g["nonpos"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_metrics_match_goldens_bit_exact(name):
    from smilecode_b200 import metrics
    g = load(name)
    flow = torch.from_numpy(g["flow"]).cuda()
    seg_m, seg_f = torch.from_numpy(g["seg_m"]).cuda(), torch.from_numpy(g["seg_f"]).cuda()
    shape = tuple(flow.shape[2:])
    warped = metrics.register_model(shape, "nearest")([seg_m, flow])          # infer.py:66,87
    assert np.array_equal(warped.cpu().numpy(), g["warped"])
    assert metrics.dice_val_VOI(warped.long(), seg_f.long()) == float(g["dsc"])   # infer.py:91
    det, nonpos = metrics.jacobian_determinant_vxm(flow)                      # infer.py:89
    assert np.array_equal(det.cpu().numpy(), g["det"])                        # float64, operation for operation
    assert int(nonpos.item()) == int(g["nonpos"])
    assert metrics.nonpositive_jacobian_fraction(flow) == int(g["nonpos"]) / det.numel()   # infer.py:90


@pytest.mark.gpu
def test_gpu_metrics_larger_random_vs_oracle():
    from smilecode_b200 import metrics
    gen = torch.Generator().manual_seed(5)
    shape = (37, 22, 45)
    coarse = torch.randn(1, 3, 6, 5, 7, generator=gen) * 4
    flow = torch.nn.functional.interpolate(coarse, size=shape, mode="trilinear", align_corners=True)
    flow = flow + 0.1 * torch.randn(flow.shape, generator=gen)
    flow[0, :, 0, 0, 0] = torch.tensor([-3.0, 0.5, 100.0])      # out-of-volume samples -> 0
    flow[0, :, 1, 1, 1] = torch.tensor([0.5, 0.5, 0.5])          # ties: round half to even
    seg = torch.randint(0, 70, (1, 2, *shape), generator=gen).float()
    ref = morc.warp_nearest(seg, flow)
    out = metrics.warp3d_nearest(seg.cuda(), flow.cuda())
    assert torch.equal(out.cpu(), ref)
    seg2 = torch.where(torch.rand(seg.shape, generator=gen) < 0.6, seg, torch.zeros_like(seg))
    labels = [1, 2, 3, 5, 8, 13, 21, 34, 55, 69]
    d_ref = morc.dice_val_VOI(seg[0, 0].long().numpy(), seg2[0, 0].long().numpy(), labels)
    assert metrics.dice_val_VOI(seg.cuda(), seg2.cuda(), labels) == d_ref
    det_ref = morc.jacobian_determinant_vxm(flow.numpy()[0])
    det, nonpos = metrics.jacobian_determinant_vxm(flow.cuda())
    assert np.array_equal(det.cpu().numpy(), det_ref)
    assert int(nonpos.item()) == int(np.sum(det_ref <= 0))


@pytest.mark.gpu
def test_gpu_metrics_errors_are_loud():
    from smilecode_b200 import metrics
    from smilecode_b200._lib import SmileError
    with pytest.raises(SmileError):
        metrics.warp3d_nearest(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 3, 4, 4, 4))          # CPU tensors
    with pytest.raises(SmileError):
        metrics.jacobian_determinant_vxm(torch.zeros(3, 1, 4, 4, device="cuda"))                  # np.gradient needs >= 2
    with pytest.raises(SmileError):
        metrics.dice_val_VOI(torch.zeros(1, 1, 2, 2, 2, device="cuda"), torch.zeros(1, 1, 2, 2, 2, device="cuda"),
                             list(range(300)))                                                     # too many labels
