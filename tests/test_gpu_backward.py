"""Backward kernels vs torch autograd of the CPU oracle (the reference gets its gradients from autograd over
the same library ops).  fp32; tolerances are relative to the gradient's max magnitude."""
import pytest
import torch

from oracle import modet_oracle as orc

pytestmark = pytest.mark.gpu


def dev(t):
    return t.detach().cuda().contiguous()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("shape,C,amp", [((6, 7, 9), 3, 2.0), ((4, 10, 33), 5, 6.0), ((2, 2, 2), 1, 0.7)])
def test_warp_backward(shape, C, amp):
    from smilecode_b200.autograd import Warp
    g = torch.Generator().manual_seed(20)
    src = torch.randn(2, C, *shape, generator=g)
    flow = torch.randn(2, 3, *shape, generator=g) * amp
    G = torch.randn(2, C, *shape, generator=g)
    s_ref, f_ref = src.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    (orc.warp_trilinear(s_ref, f_ref) * G).sum().backward()
    s_d, f_d = dev(src).requires_grad_(True), dev(flow).requires_grad_(True)
    (Warp.apply(s_d, f_d) * dev(G)).sum().backward()
    assert rel(s_d.grad.cpu(), s_ref.grad) <= 1e-5
    assert rel(f_d.grad.cpu(), f_ref.grad) <= 2e-4      # weights differences cancel: conditioning ~ |src| / |d src|


@pytest.mark.parametrize("shape,C", [((3, 4, 5), 3), ((10, 12, 10), 2), ((2, 2, 2), 24)])
def test_upsample_backward(shape, C):
    from smilecode_b200.autograd import Upsample2x
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, C, *shape, generator=g)
    G = torch.randn(2, C, *[2 * s for s in shape], generator=g)
    xr = x.clone().requires_grad_(True)
    (2.0 * orc.upsample2x_trilinear(xr) * G).sum().backward()
    xd = dev(x).requires_grad_(True)
    (Upsample2x.apply(xd, 2.0) * dev(G)).sum().backward()
    assert rel(xd.grad.cpu(), xr.grad) <= 1e-5


@pytest.mark.parametrize("shape,heads,hd,scale", [((1, 5, 6, 7), 4, 6, 1.0), ((2, 3, 8, 32), 1, 6, 1.0), ((1, 2, 2, 2), 8, 6, 0.4),
                                                  ((1, 4, 5, 6), 6, 6, 1.0)])
def test_attention_backward(shape, heads, hd, scale):
    from smilecode_b200.autograd import Attention
    B, D, H, W = shape
    g = torch.Generator().manual_seed(22)
    q = torch.randn(B, D, H, W, heads * hd, generator=g)
    k = torch.randn(B, D, H, W, heads * hd, generator=g)
    rpb = torch.randn(heads, 3, 3, 3, generator=g) * 0.5
    G = torch.randn(B, 3 * heads, D, H, W, generator=g)
    qr, kr, rr = (t.clone().requires_grad_(True) for t in (q, k, rpb))
    (orc.modet_attention(qr, kr, rr, heads, scale) * G).sum().backward()
    qd, kd, rd = (dev(t).requires_grad_(True) for t in (q, k, rpb))
    (Attention.apply(qd, kd, rd, heads, scale) * dev(G)).sum().backward()
    assert rel(qd.grad.cpu(), qr.grad) <= 2e-5
    assert rel(kd.grad.cpu(), kr.grad) <= 2e-5
    assert rel(rd.grad.cpu(), rr.grad) <= 2e-5


@pytest.mark.parametrize("cin,c,shape", [(8, 6, (5, 6, 7)), (16, 6, (3, 4, 33)), (128, 48, (2, 3, 2)), (32, 12, (4, 4, 4))])
def test_projection_backward(cin, c, shape):
    from smilecode_b200.autograd import ProjLN
    g = torch.Generator().manual_seed(23)
    x = torch.randn(2, cin, *shape, generator=g)
    sd = {"p.proj.weight": torch.randn(c, cin, generator=g) / cin ** 0.5, "p.proj.bias": torch.randn(c, generator=g) * 0.1,
          "p.norm.weight": torch.rand(c, generator=g) + 0.5, "p.norm.bias": torch.randn(c, generator=g) * 0.1}
    G = torch.randn(2, *shape, c, generator=g)
    xr = x.clone().requires_grad_(True)
    sdr = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
    (orc.projection(xr, sdr, "p") * G).sum().backward()
    xd = dev(x).requires_grad_(True)
    pd = {k_: dev(v).requires_grad_(True) for k_, v in sd.items()}
    out = ProjLN.apply(xd, pd["p.proj.weight"], pd["p.proj.bias"], pd["p.norm.weight"], pd["p.norm.bias"], 1e-5)
    (out * dev(G)).sum().backward()
    assert rel(xd.grad.cpu(), xr.grad) <= 5e-5
    for k_ in sd:
        assert rel(pd[k_].grad.cpu(), sdr[k_].grad) <= 5e-5, k_


@pytest.mark.parametrize("F,shape", [(8, (4, 6, 4)), (2, (9, 5, 7))])
def test_cwm_fuse_backward(F, shape):
    from smilecode_b200.autograd import CwmFuse
    g = torch.Generator().manual_seed(24)
    u = torch.randn(2, 3 * F, *shape, generator=g)
    lg = torch.randn(2, F, *shape, generator=g)
    G = torch.randn(2, 3, *shape, generator=g)

    def ref(u_, lg_):
        p = lg_.softmax(1)
        return 2 * sum(u_[:, 3 * f:3 * f + 3] * p[:, f:f + 1] for f in range(F))
    ur, lr = u.clone().requires_grad_(True), lg.clone().requires_grad_(True)
    (ref(ur, lr) * G).sum().backward()
    ud, ld = dev(u).requires_grad_(True), dev(lg).requires_grad_(True)
    (CwmFuse.apply(ud, ld) * dev(G)).sum().backward()
    assert rel(ud.grad.cpu(), ur.grad) <= 1e-5
    assert rel(ld.grad.cpu(), lr.grad) <= 1e-5
