"""Skinny head GEMMs (hid -> classes): dedicated streaming kernels behind gda_gemm_f32."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,k,c", [(5000, 128, 5), (100_000, 128, 5), (4096, 256, 2), (3001, 64, 16)])
def test_forward_dx_dw(rows, k, c):
    from pygda_b200 import ops
    x, w, g = torch.randn(rows, k), torch.randn(c, k), torch.randn(rows, c)
    xc, wc, gc = x.cuda(), w.cuda(), g.cuda()
    assert_close(ops.gemm(xc, wc, trans_b=True), x.double() @ w.double().t(), 1e-5, "fwd")
    assert_close(ops.gemm(gc, wc), g.double() @ w.double(), 1e-5, "dx")
    assert_close(ops.gemm(gc, xc, trans_a=True), g.double().t() @ x.double(), 1e-5, "dw")


def test_cls_layer_autograd_at_benchmark_width():
    """PropGCNConv(128 -> 5) with one propagation step: the classifier of A2GNNBase."""
    from pygda_b200.nn import PropGCNConv
    from pygda_b200.synthetic import powerlaw_edge_index
    from oracle import nn as ONN
    n = 20000
    ei = powerlaw_edge_index(n, 200000, seed=3)
    torch.manual_seed(0)
    ref = ONN.PropGCNConv(128, 5)
    conv = PropGCNConv(128, 5).cuda()
    conv.load_state_dict(ref.state_dict())
    x = torch.randn(n, 128)
    coef = torch.randn(n, 5)
    xr = x.clone().requires_grad_(True)
    (ref(xr, ei, 1) * coef).sum().backward()
    xg = x.cuda().requires_grad_(True)
    y = conv(xg, ei.cuda(), 1)
    (y * coef.cuda()).sum().backward()
    assert_close(y, ref(x, ei, 1), 1e-4, "fwd")
    assert_close(conv.lin.weight.grad, ref.lin.weight.grad, 1e-4, "dW")
    assert_close(xg.grad, xr.grad, 1e-4, "dx")


@pytest.mark.parametrize("ta,tb,m,n,k", [(False, True, 200_000, 40, 256), (False, False, 200_000, 256, 40),
                                         (False, False, 150_000, 33, 128), (True, False, 300, 40, 100_000)])
def test_narrow_widths_are_padded_onto_the_tensor_cores_exactly(ta, tb, m, n, k):
    """Widths between the skinny kernels (<= 16) and the tcgen05 kernel (>= 64) -- the 40-wide domain MLP of
    UDAGCN / AdaGCN (pygda/nn/udagcn_base.py:157-162) -- are zero-padded to 64 (ops.mm): same values."""
    from pygda_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn((k, m) if ta else (m, k), generator=g).cuda()
    b = torch.randn((n, k) if tb else (k, n), generator=g).cuda()
    out, _, _ = ops.mm(a, b, trans_a=ta, trans_b=tb)
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    assert out.shape == (m, n) and out.is_contiguous()
    assert_close(out, ref, 3e-5, "padded tensor-core product")


def test_linear_256_to_40_forward_and_backward():
    from pygda_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(120_000, 256, device="cuda", requires_grad=True)
    w = (torch.randn(40, 256, device="cuda") * 0.1).requires_grad_(True)
    b = torch.randn(40, device="cuda", requires_grad=True)
    y = ops.linear(x, w, b)
    go = torch.randn_like(y)
    y.backward(go)
    xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yr = xr @ wr.t() + br
    yr.backward(go.double())
    assert_close(y, yr, 3e-5, "forward")
    assert_close(x.grad, xr.grad, 3e-5, "dX (reduction width 40)")
    assert_close(w.grad, wr.grad, 1e-4, "dW")
    assert_close(b.grad, br.grad, 1e-4, "db")
