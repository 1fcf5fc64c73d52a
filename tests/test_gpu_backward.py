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


@pytest.mark.parametrize("cin,cout,shape,pool", [(4, 8, (6, 8, 34), True), (1, 4, (5, 6, 7), False), (16, 16, (4, 6, 20), True),
                                                 (64, 128, (4, 6, 10), False), (6, 12, (3, 5, 33), False)])
def test_conv_in_lrelu_backward(cin, cout, shape, pool):
    from smilecode_b200.autograd import ConvINLReLU
    g = torch.Generator().manual_seed(25)
    x = torch.randn(2, cin, *shape, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    G = torch.randn(2, cout, *shape, generator=g)
    Gp = torch.randn(2, cout, *[s // 2 for s in shape], generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    act = orc.lrelu(orc.instance_norm(orc.conv3(xr, wr, br)))
    loss = (act * G).sum()
    if pool:
        loss = loss + (torch.nn.functional.avg_pool3d(act, 2) * Gp).sum()
    loss.backward()
    xd, wd, bd = (dev(t).requires_grad_(True) for t in (x, w, b))
    if pool:
        a, p = ConvINLReLU.apply(xd, wd, bd, True)
        ((a * dev(G)).sum() + (p * dev(Gp)).sum()).backward()
    else:
        (ConvINLReLU.apply(xd, wd, bd, False) * dev(G)).sum().backward()
    assert rel(xd.grad.cpu(), xr.grad) <= 1e-4
    assert rel(wd.grad.cpu(), wr.grad) <= 1e-4
    assert bd.grad.abs().max() <= 1e-3 * wd.grad.abs().max()          # bias before InstanceNorm has zero gradient


def test_conv_lrelu_and_plain_conv_backward():
    from smilecode_b200.autograd import Conv, ConvLReLU
    g = torch.Generator().manual_seed(26)
    x = torch.randn(2, 3, 5, 7, 36, generator=g)
    w = torch.randn(5, 3, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(5, generator=g) * 0.1
    G = torch.randn(2, 5, 5, 7, 36, generator=g)
    for fn, ref in ((ConvLReLU, lambda a, c, e: orc.lrelu(orc.conv3(a, c, e))), (Conv, orc.conv3)):
        xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
        (ref(xr, wr, br) * G).sum().backward()
        xd, wd, bd = (dev(t).requires_grad_(True) for t in (x, w, b))
        (fn.apply(xd, wd, bd) * dev(G)).sum().backward()
        assert rel(xd.grad.cpu(), xr.grad) <= 2e-5
        assert rel(wd.grad.cpu(), wr.grad) <= 2e-5
        assert rel(bd.grad.cpu(), br.grad) <= 2e-5


@pytest.mark.parametrize("cin,cout,shape,batch", [(8, 8, (6, 12, 40), 2), (4, 8, (5, 7, 16), 1), (6, 12, (4, 9, 20), 2),
                                                  (16, 16, (3, 6, 10), 2), (32, 64, (5, 6, 10), 2), (128, 128, (2, 3, 5), 2),
                                                  (3, 5, (4, 5, 7), 3), (20, 70, (3, 4, 6), 1)])
def test_conv_weight_gradient_on_tensor_cores(cin, cout, shape, batch):
    """smile_conv3d_wgrad_bf16 (tcgen05, im2col built in shared memory, K = positions): against the fp32 weight gradient of
    the SAME bf16-rounded operands (tight: only the accumulation order differs) and against the fp32 gradient of the
    original operands (bf16 rounding: ~3e-3 relative).  Row-crossing position groups (W not a multiple of 8), partial
    channel tiles, several output-channel tiles, batch > 1, tails shorter than a chunk."""
    from smilecode_b200 import ops
    g = torch.Generator().manual_seed(31)
    x = torch.randn(batch, cin, *shape, generator=g)
    gy = torch.randn(batch, cout, *shape, generator=g)
    w = torch.zeros(cout, cin, 3, 3, 3)

    def ref_grads(xx, gg):
        wr = w.clone().double().requires_grad_(True)
        br = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
        (torch.nn.functional.conv3d(xx.double(), wr, br, padding=1) * gg.double()).sum().backward()
        return wr.grad.float(), br.grad.float()

    with ops.conv_precision("bf16"):
        _, dw, db = ops.conv3d_bwd(dev(gy), dev(x), dev(w), need_x=False)
    dw_r, db_r = ref_grads(x, gy)
    dw_q, _ = ref_grads(x.bfloat16().float(), gy.bfloat16().float())
    assert rel(dw.cpu(), dw_q) <= 2e-5, (cin, cout, shape)
    assert rel(dw.cpu(), dw_r) <= 1e-2
    assert rel(db.cpu(), db_r) <= 1e-5


def test_losses_backward():
    from smilecode_b200.autograd import Grad3dLoss, NCCLoss
    from smilecode_b200.synth import make_pair
    moving, fixed = make_pair((20, 24, 18), batch=2, seed=5)
    mr = moving.clone().requires_grad_(True)
    (3.0 * orc.ncc_vxm(mr, fixed)).backward()
    md = dev(moving).requires_grad_(True)
    (3.0 * NCCLoss.apply(md, dev(fixed), 9)).backward()
    assert rel(md.grad.cpu(), mr.grad) <= 2e-3           # cc has cancellation in flat regions (var ~ 1e-5)
    flow = torch.randn(2, 3, 7, 9, 11, generator=torch.Generator().manual_seed(6))
    fr = flow.clone().requires_grad_(True)
    (0.5 * orc.grad3d_l2(fr)).backward()
    fd = dev(flow).requires_grad_(True)
    (0.5 * Grad3dLoss.apply(fd)).backward()
    assert rel(fd.grad.cpu(), fr.grad) <= 1e-5


def test_adam_amsgrad_matches_torch():
    from smilecode_b200 import ops
    g = torch.Generator().manual_seed(27)
    p0 = torch.randn(1000, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, amsgrad=True)
    p = dev(p0)
    m, v, vm = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(1000, generator=g) * (0.1 if step == 4 else 1.0)
        ref.grad = grad.clone()
        opt.step()
        ops.adam_amsgrad_step(p, dev(grad), m, v, vm, 1e-2, 0.9, 0.999, 1e-8, step)
    assert (p.cpu() - ref.detach()).abs().max() <= 1e-6


def test_end_to_end_training_gradients_match_reference_autograd():
    """One training step's gradients (NCC + Grad3d loss, train.py:126-132) for every parameter vs autograd of the
    CPU oracle of the whole model."""
    from smilecode_b200 import losses, models
    from smilecode_b200.synth import make_pair
    shape, heads = (16, 32, 32), [8, 4, 2, 1, 1]
    sd = orc.synth_state_dict(seed=1234, num_heads=heads)
    moving, fixed = make_pair(shape, batch=1, seed=24)
    # reference gradients on the CPU
    sd_r = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("grid")) for k, v in sd.items()}
    y, flow = orc.modet_forward(moving, fixed, sd_r, num_heads=heads, scale=1.0, library_ops=True)
    loss_r = orc.ncc_vxm(y, fixed) + orc.grad3d_l2(flow)
    loss_r.backward()
    # ours
    model = models.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().train()
    y_d, flow_d = model(moving.cuda(), fixed.cuda())
    loss = losses.NCC_vxm()(y_d, fixed.cuda()) + losses.Grad3d(penalty="l2")(flow_d, fixed.cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_r.detach())) <= 1e-4 * max(1.0, abs(float(loss_r)))
    worst = {}
    for name, p in model.named_parameters():
        gr = sd_r[name].grad
        assert gr is not None and p.grad is not None, name
        worst[name] = float((p.grad.cpu() - gr).abs().max() / gr.abs().max().clamp_min(1e-12))
    # a bias in front of an InstanceNorm has exactly zero gradient (the norm removes the mean): only noise there
    live = {k: v for k, v in worst.items() if not (k.endswith("main.bias") and ".conv.2." not in k and "conv0.0" not in k)}
    print("worst relative gradient errors:", sorted(live.items(), key=lambda kv: -kv[1])[:5])
    bad = {k: v for k, v in live.items() if v > 1e-3}
    assert not bad, bad


def test_trainer_steps_reduce_the_loss_and_match_torch_adam():
    from smilecode_b200 import models
    from smilecode_b200.synth import make_pair, randomize_weights
    from smilecode_b200.train import Trainer
    shape, heads = (16, 32, 32), [8, 4, 2, 1, 1]
    model = models.ModeT(shape, head_dim=6, num_heads=heads, scale=1)
    randomize_weights(model, seed=3)
    model = model.cuda()
    moving, fixed = (t.cuda() for t in make_pair(shape, batch=1, seed=9))
    # twin updated by torch.optim.Adam from the same gradients (first step)
    import copy
    twin = copy.deepcopy(model)
    tr = Trainer(model, lr=1e-3)
    losses_seen = []
    for it in range(6):
        if it == 0:
            opt = torch.optim.Adam(twin.parameters(), lr=1e-3, amsgrad=True)
        loss, _, _ = tr.step(moving, fixed)
        losses_seen.append(float(loss))
        if it == 0:
            for p, q in zip(model.parameters(), twin.parameters()):
                q.grad = p.grad.clone()
            opt.step()
            for (n, p), q in zip(model.named_parameters(), twin.parameters()):
                assert (p.data - q.data).abs().max() <= 1e-6, n
    assert losses_seen[-1] < losses_seen[0], losses_seen
