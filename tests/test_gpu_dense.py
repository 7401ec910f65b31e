"""Dense GEMM, elementwise, loss, MMD, pooling and Adam kernels vs plain torch fp32/fp64."""
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close, load_golden
from oracle import mmd as OM
from oracle import pyg_ops as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (37, 5, 128), (300, 128, 6775), (1000, 2, 128),
                                   (128, 677, 9000), (257, 129, 65), (4096, 128, 128)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_all_layouts(m, n, k, ta, tb):
    from pygda_b200 import ops
    a = torch.randn((k, m) if ta else (m, k))
    b = torch.randn((n, k) if tb else (k, n))
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    out = ops.gemm(a.cuda(), b.cuda(), trans_a=ta, trans_b=tb)
    assert_close(out, ref, 1e-5, f"gemm {m}x{n}x{k} ta={ta} tb={tb}")


def test_gemm_alpha_beta_and_strided_output():
    from pygda_b200 import ops
    a, b = torch.randn(70, 33), torch.randn(33, 20)
    c0 = torch.randn(70, 20)
    out = ops.gemm(a.cuda(), b.cuda(), alpha=0.5, beta=2.0, out=c0.cuda().clone())
    assert_close(out, 0.5 * a @ b + 2.0 * c0, 1e-5, "alpha/beta")


def test_linear_and_graphconv_autograd_match_oracle_golden():
    from pygda_b200.nn import PropGCNConv
    g = load_golden("prop_gcn_conv")
    conv = PropGCNConv(12, 8).cuda()
    conv.load_state_dict(g["state"])
    ei = g["edge_index"].cuda()
    for k in (0, 1, 3):
        x = g["x"].cuda().requires_grad_(True)
        conv.zero_grad()
        y = conv(x, ei, k)
        (y * torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
        assert_close(y, g["out"][k], 1e-5, f"out k={k}")
        assert_close(conv.lin.weight.grad, g["grad_w"][k], 1e-4, f"grad_w k={k}")
        assert_close(x.grad, g["grad_x"][k], 1e-4, f"grad_x k={k}")


def test_bias_act_dropout_fwd_bwd():
    from pygda_b200 import ops
    x = torch.randn(500, 64)
    xg = x.cuda().requires_grad_(True)
    y = ops.ActDropoutFn.apply(xg, 1, 0.0, 0)
    assert_close(y, torch.relu(x), 1e-7, "relu")
    y = ops.ActDropoutFn.apply(xg, 1, 0.3, 42)
    g = torch.randn(500, 64)
    y.backward(g.cuda())
    keep = (y != 0).cpu()
    ref = torch.where(keep, g / 0.7, torch.zeros_like(g))
    assert_close(xg.grad, ref, 1e-6, "dropout bwd regenerates the mask")
    frac = keep.float().sum() / (x > 0).float().sum()
    assert 0.66 < float(frac) < 0.74


def test_softmax_ce_and_domain_labels():
    from pygda_b200 import ops
    z = torch.randn(1000, 5) * 3
    y = torch.randint(5, (1000,))
    zr = z.clone().requires_grad_(True)
    ref = F.nll_loss(F.log_softmax(zr, dim=1), y)
    (ref * 1.7).backward()
    zg = z.cuda().requires_grad_(True)
    loss = ops.softmax_cross_entropy(zg, y.cuda())
    (loss * 1.7).backward()
    assert_close(loss, ref, 1e-5, "ce")
    assert_close(zg.grad, zr.grad, 1e-5, "ce grad")
    d = torch.randn(300, 2)
    lab = torch.tensor([0] * 120 + [1] * 180)
    assert_close(ops.domain_cross_entropy(d.cuda(), 120), F.cross_entropy(d, lab), 1e-5, "domain ce")


@pytest.mark.parametrize("c", [64, 65, 100, 1000])
def test_softmax_ce_many_classes_and_bad_labels(c):
    """More than 64 classes take the warp-per-row kernel (the reference supports any num_classes); a label
    outside [0, C) raises like F.nll_loss does instead of indexing out of bounds."""
    from pygda_b200 import ops
    g = torch.Generator().manual_seed(c)
    z = torch.randn(513, c, generator=g) * 3
    y = torch.randint(c, (513,), generator=g)
    zr = z.clone().requires_grad_(True)
    ref = F.nll_loss(F.log_softmax(zr, dim=1), y)
    ref.backward()
    zg = z.cuda().requires_grad_(True)
    loss = ops.softmax_cross_entropy(zg, y.cuda())
    loss.backward()
    assert_close(loss, ref, 1e-5, f"ce, C={c}")
    assert_close(zg.grad, zr.grad, 1e-5, f"ce grad, C={c}")
    for bad in (-100, -1, c):
        yb = y.clone()
        yb[7] = bad
        with pytest.raises(IndexError):
            ops.softmax_cross_entropy(z.cuda(), yb.cuda())


def test_softmax_entropy():
    from pygda_b200 import ops
    z = torch.randn(700, 6) * 4
    zr = z.clone().requires_grad_(True)
    p = torch.clamp(F.softmax(zr, dim=-1), min=1e-9, max=1.0)
    ref = torch.mean(torch.sum(-p * torch.log(p), dim=-1))
    ref.backward()
    zg = z.cuda().requires_grad_(True)
    out = ops.softmax_entropy(zg)
    out.backward()
    assert_close(out, ref, 1e-5, "entropy")
    assert_close(zg.grad, zr.grad, 1e-4, "entropy grad")


def test_mmd_golden_from_reference_file():
    from pygda_b200.utils import MMD, get_MMD
    g = load_golden("mmd")
    s, t = g["source"].cuda().requires_grad_(True), g["target"].cuda().requires_grad_(True)
    loss = MMD(s, t, indices=(g["source_idx"], g["target_idx"]))
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "mmd loss")
    assert_close(s.grad, g["grad_source"], 1e-4, "mmd grad source")
    assert_close(t.grad, g["grad_target"], 1e-4, "mmd grad target")
    assert_close(get_MMD(g["source"][:50].cuda(), g["target"][:50].cuda()), g["get_mmd_full"], 1e-4, "get_MMD")


def test_mmd_default_size_vs_oracle_and_index_draws():
    from pygda_b200.utils import MMD
    torch.manual_seed(3)
    s, t = torch.randn(3000, 128) * 0.3, torch.randn(2500, 128) * 0.4 + 0.1
    torch.manual_seed(11)
    idx = OM.draw_mmd_indices(3000, 2500)
    sr, tr = s.clone().requires_grad_(True), t.clone().requires_grad_(True)
    ref = OM.MMD(sr, tr, indices=idx, sqdist=lambda z: OM.pairwise_sqdist_blocked(z, 128))
    ref.backward()
    sg, tg = s.cuda().requires_grad_(True), t.cuda().requires_grad_(True)
    torch.manual_seed(11)
    out = MMD(sg, tg)                       # draws its own indices from the CPU generator
    out.backward()
    assert_close(out, ref, 1e-4, "mmd n=2000")
    assert_close(sg.grad, sr.grad, 1e-4, "grad source")
    assert_close(tg.grad, tr.grad, 1e-4, "grad target")


def test_global_mean_pool():
    from pygda_b200 import ops
    x = torch.randn(1000, 128)
    sizes = torch.randint(1, 40, (60,))
    sizes[-1] = 1000 - sizes[:-1].sum() if sizes[:-1].sum() < 1000 else 1
    batch = torch.repeat_interleave(torch.arange(60), sizes)[:1000]
    x = x[: batch.numel()]
    xr = x.clone().requires_grad_(True)
    ref = P.global_mean_pool(xr, batch)
    coef = torch.randn_like(ref)
    (ref * coef).sum().backward()
    xg = x.cuda().requires_grad_(True)
    out = ops.global_mean_pool(xg, batch.cuda())
    (out * coef.cuda()).sum().backward()
    assert_close(out, ref, 1e-5, "pool")
    assert_close(xg.grad, xr.grad, 1e-5, "pool grad")


def test_adam_matches_torch_optim():
    from pygda_b200.optim import Adam
    torch.manual_seed(0)
    shapes = [(128, 300), (128,), (5, 128), (5,)]
    ps = [torch.randn(s) for s in shapes]
    ref_p = [p.clone().requires_grad_(True) for p in ps]
    my_p = [p.clone().cuda().requires_grad_(True) for p in ps]
    ref_opt = torch.optim.Adam(ref_p, lr=0.01, weight_decay=0.005)
    my_opt = Adam(my_p, lr=0.01, weight_decay=0.005)
    for step in range(5):
        gs = [torch.randn(s) for s in shapes]
        for p, q, g in zip(ref_p, my_p, gs):
            p.grad = g.clone()
            q.grad = g.clone().cuda()
        ref_opt.step()
        my_opt.step()
    for p, q in zip(ref_p, my_p):
        assert_close(q, p, 1e-5, "adam param")


def test_grad_reverse_golden():
    from pygda_b200.nn import GradReverse
    g = load_golden("grad_reverse")
    x = g["x"].cuda().requires_grad_(True)
    y = GradReverse.apply(x, g["alpha"])
    y.backward(torch.ones_like(y) * 2.0)
    assert torch.equal(y.cpu(), g["y"]) and torch.allclose(x.grad.cpu(), g["grad"])
