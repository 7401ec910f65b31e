"""tcgen05 split-bf16 GEMM vs fp64: all operand majors, tails in M/N/K, split-K, precision."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


def _run(m, n, k, ta, tb, seed=0):
    from pygda_b200 import ops
    g = torch.Generator().manual_seed(seed)
    a = torch.randn((k, m) if ta else (m, k), generator=g)
    b = torch.randn((n, k) if tb else (k, n), generator=g)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    sa, sb = ops.Split(a.cuda()), ops.Split(b.cuda())
    out = ops.gemm_split(sa, sb, ta, tb, m, n, k)
    torch.cuda.synchronize()
    return out, ref


def test_split_reconstructs_to_2pow17():
    from pygda_b200 import ops
    x = torch.randn(300, 77) * torch.logspace(-3, 3, 77)
    sp = ops.Split(x.cuda())
    assert sp.ld == 80 and sp.hi.shape == (300, 80)
    rec = sp.hi.float() + sp.lo.float()
    assert torch.all(rec[:, 77:] == 0)
    err = (rec[:, :77].cpu() - x).abs() / x.abs().clamp(min=1e-30)
    assert float(err.max()) < 2.0 ** -16


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_single_tile_all_majors(ta, tb):
    out, ref = _run(128, 128, 64, ta, tb)
    assert_close(out, ref, 3e-5, f"128x128x64 ta={ta} tb={tb}")


@pytest.mark.parametrize("m,n,k,ta,tb", [
    (1000, 128, 6775, False, True),      # layer-1 forward shape (K tail 6775 % 64 != 0, M tail)
    (128, 6775, 5000, True, False),      # layer-1 weight gradient: split-K, N tail, MN-major both
    (777, 128, 128, False, False),       # hidden dX = G W (B MN-major)
    (128, 128, 20000, True, False),      # hidden dW: single tile, deep split-K
    (300, 200, 333, True, True),
    (129, 65, 64, False, True),
])
def test_shapes_tails_and_splitk(m, n, k, ta, tb):
    out, ref = _run(m, n, k, ta, tb, seed=m + n + k)
    assert_close(out, ref, 3e-5, f"{m}x{n}x{k} ta={ta} tb={tb}")


def test_beats_plain_bf16_and_tf32_error():
    """The 3-term split keeps ~fp32 accuracy: error far below a single bf16 product."""
    out, ref = _run(512, 128, 4096, False, True, seed=5)
    err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


def test_mm_dispatch_and_autograd_large_linear():
    """GraphConvFn k=0 at a tensor-core-eligible size: forward, dW and dx vs fp64."""
    from pygda_b200 import ops
    n, f, h = 4096, 512, 128
    x, w, b = torch.randn(n, f), torch.randn(h, f) * 0.05, torch.randn(h) * 0.1
    assert ops.tc_eligible(n, h, f)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    coef = torch.randn(n, h)
    ((xr @ wr.t() + b.double()) * coef.double()).sum().backward()
    xg, wg, bg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.graph_conv(xg, wg, bg, None, 0)
    (y * coef.cuda()).sum().backward()
    assert_close(y, (x.double() @ w.double().t() + b.double()), 3e-5, "fwd")
    assert_close(wg.grad, wr.grad, 3e-5, "dW")
    assert_close(xg.grad, xr.grad, 3e-5, "dx")
    assert_close(bg.grad, coef.sum(0), 1e-5, "db")
