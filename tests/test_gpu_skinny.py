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
