"""Aggregation kernel vs the oracle's index_select -> mul -> scatter_add_."""
import pytest
import torch

from conftest import assert_close
from oracle import pyg_ops as P

pytestmark = pytest.mark.gpu


def powerlaw(n, e, seed):
    from pygda_b200.synthetic import powerlaw_edge_index
    return powerlaw_edge_index(n, e, seed=seed, offset=2.0)


def _setup(n, e, seed, directed=False):
    from pygda_b200.graph import Graph
    ei = powerlaw(n, e, seed)
    if directed:
        ei = ei[:, : e // 2]
    gr = Graph(ei.cuda(), n)
    ref_ei, ref_w = P.gcn_norm_by_col(ei, None, n)
    return gr, ref_ei, ref_w


@pytest.mark.parametrize("h", [1, 2, 5, 7, 16, 64, 100, 128, 256, 300])
def test_widths_forward_and_transpose(h):
    from pygda_b200 import ops
    n = 3000
    gr, ref_ei, ref_w = _setup(n, 40000, 1, directed=True)
    x = torch.randn(n, h)
    y = ops.spmm(gr, x.cuda())
    assert_close(y, P.propagate(ref_ei, x, ref_w), 1e-5, f"A x, H={h}")
    yt = ops.spmm(gr, x.cuda(), transpose=True)
    assert_close(yt, P.propagate(ref_ei.flip(0), x, ref_w), 1e-5, f"A^T x, H={h}")


def test_hub_rows_and_determinism():
    from pygda_b200 import ops
    n = 20000
    gr, ref_ei, ref_w = _setup(n, 400000, 2)
    assert gr.num_long_rows > 0
    x = torch.randn(n, 128)
    xc = x.cuda()
    y1 = ops.spmm(gr, xc)
    y2 = ops.spmm(gr, xc)
    assert torch.equal(y1, y2)                      # counters reset, fixed summation order
    assert_close(y1, P.propagate(ref_ei, x, ref_w), 1e-5, "hub graph")


def test_k_steps_match_repeated_propagate():
    from pygda_b200 import ops
    n = 4000
    gr, ref_ei, ref_w = _setup(n, 50000, 3)
    x = torch.randn(n, 128)
    ref = x
    for _ in range(10):
        ref = P.propagate(ref_ei, ref, ref_w)
    assert_close(ops.spmm_k(gr, x.cuda(), 10), ref, 1e-4, "A^10 x")


def test_epilogue_bias_relu_and_dropout_statistics():
    from pygda_b200 import ops
    n, h = 3000, 128
    gr, ref_ei, ref_w = _setup(n, 30000, 4)
    x, b = torch.randn(n, h), torch.randn(h)
    ref = torch.relu(P.propagate(ref_ei, x, ref_w) + b)
    y = ops.spmm(gr, x.cuda(), bias=b.cuda(), relu=True)
    assert_close(y, ref, 1e-5, "bias+relu epilogue")
    yd = ops.spmm(gr, x.cuda(), bias=b.cuda(), relu=True, dropout_p=0.5, seed=123).cpu()
    kept = yd != 0
    assert torch.allclose(yd[kept], (ref * 2.0)[kept], rtol=1e-4, atol=1e-5)
    frac = kept.float().sum() / (ref != 0).float().sum()
    assert 0.48 < float(frac) < 0.52
    yd2 = ops.spmm(gr, x.cuda(), bias=b.cuda(), relu=True, dropout_p=0.5, seed=123).cpu()
    assert torch.equal(yd, yd2)                      # same seed -> same mask


def test_bf16_features_fp32_accumulate():
    from pygda_b200 import ops
    n, h = 3000, 256
    gr, ref_ei, ref_w = _setup(n, 40000, 5)
    x = torch.randn(n, h).bfloat16()
    ref = P.propagate(ref_ei, x.float(), ref_w)
    y = ops.spmm(gr, x.cuda())
    assert y.dtype == torch.bfloat16
    assert_close(y.float(), ref, 1e-2, "bf16 spmm")   # output rounding to bf16 (2^-9)


def test_empty_graph_and_isolated_nodes():
    from pygda_b200 import ops
    from pygda_b200.graph import Graph
    n = 10
    gr = Graph(torch.zeros(2, 0, dtype=torch.long).cuda(), n)
    x = torch.randn(n, 8)
    assert_close(ops.spmm(gr, x.cuda()), x, 1e-6, "self loops only")   # A_hat = I


def test_autograd_propagate():
    from pygda_b200 import ops
    n = 2000
    gr, ref_ei, ref_w = _setup(n, 20000, 6, directed=True)
    x = torch.randn(n, 32)
    xr = x.clone().requires_grad_(True)
    ref = x
    out = xr
    for _ in range(3):
        out = P.propagate(ref_ei, out, ref_w)
    coef = torch.randn(n, 32)
    (out * coef).sum().backward()
    xg = x.cuda().requires_grad_(True)
    yg = ops.PropagateFn.apply(xg, gr, 3)
    (yg * coef.cuda()).sum().backward()
    assert_close(yg, out, 1e-4, "fwd")
    assert_close(xg.grad, xr.grad, 1e-4, "bwd")


# ---- factored chain for unit-weight graphs (spmm.cu: k_spmm_unw; A_hat^k = D S (D^2 S)^(k-1) D, no per-edge weights) ----
@pytest.mark.parametrize("k", [3, 4, 10])
@pytest.mark.parametrize("transpose", [False, True])
def test_unit_weight_chain_matches_the_oracle_and_the_weighted_kernel(k, transpose):
    from pygda_b200 import ops
    n = 6000
    gr, ref_ei, ref_w = _setup(n, 90000, 7, directed=True)          # directed: A_hat != A_hat^T
    assert ops.unit_weight_chain(gr, transpose, 128, 1, k)
    x = torch.randn(n, 128)
    ei = ref_ei.flip(0) if transpose else ref_ei
    ref = x
    for _ in range(k):
        ref = P.propagate(ei, ref, ref_w)
    y = ops.spmm_k(gr, x.cuda(), k, transpose=transpose)
    assert_close(y, ref, 2e-5, f"factored A^{k} x vs the oracle")
    w = x.cuda()
    for _ in range(k):
        w = ops.spmm(gr, w, transpose=transpose)                    # one step at a time: the weighted kernel
    assert_close(y, w, 1e-5, "factored vs weighted kernel")
    assert torch.equal(y, ops.spmm_k(gr, x.cuda(), k, transpose=transpose))      # deterministic


def test_unit_weight_chain_hub_rows_epilogue_and_pairs():
    from pygda_b200 import ops
    n, h, k = 20000, 128, 5
    gr, ref_ei, ref_w = _setup(n, 400000, 2)
    assert gr.num_long_rows > 0 and ops.unit_weight_chain(gr, False, h, 2, k)
    x, b = torch.randn(2 * n, h), torch.randn(h)
    refs = []
    for half in (x[:n], x[n:]):
        r = half
        for _ in range(k):
            r = P.propagate(ref_ei, r, ref_w)
        refs.append(torch.relu(r + b))
    y = ops.spmm_k(gr, x.cuda(), k, bias=b.cuda(), relu=True, nb=2)
    assert_close(y[:n], refs[0], 2e-5, "stacked pair, first half")
    assert_close(y[n:], refs[1], 2e-5, "stacked pair, second half")
    y1 = ops.spmm_k(gr, x[:n].cuda().contiguous(), k, bias=b.cuda(), relu=True)
    assert torch.equal(y1, y[:n])                                   # nb = 2 == two nb = 1 chains, bit for bit
    # dropout mask of the last step: that of the weighted kernel for the same seed
    yd = ops.spmm_k(gr, x[:n].cuda().contiguous(), k, bias=b.cuda(), relu=True, dropout_p=0.5, seed=99)
    wd = x[:n].cuda()
    for i in range(k):
        last = i == k - 1
        wd = ops.spmm(gr, wd, bias=b.cuda() if last else None, relu=last, dropout_p=0.5 if last else 0.0, seed=99)
    assert_close(yd, wd, 1e-5, "dropout epilogue")


def test_weighted_and_improved_graphs_do_not_take_the_factored_chain():
    from pygda_b200 import ops
    from pygda_b200.graph import Graph, SELF_LOOPS, NORM_SYM_COL, IMPROVED
    n = 3000
    ei = powerlaw(n, 30000, 5).cuda()
    w = torch.rand(ei.size(1), device="cuda") + 0.5
    assert not ops.unit_weight_chain(Graph(ei, n, w), False, 128, 1, 10)
    assert not ops.unit_weight_chain(Graph(ei, n, None, SELF_LOOPS | NORM_SYM_COL | IMPROVED), False, 128, 1, 10)
    assert ops.unit_weight_chain(Graph(ei, n), False, 128, 1, 10)
    assert not ops.unit_weight_chain(Graph(ei, n), False, 64, 1, 10)        # other widths: weighted kernels
    x = torch.randn(n, 128)
    ref_ei, ref_w = P.gcn_norm_by_col(ei.cpu(), w.cpu(), n)
    ref = x
    for _ in range(4):
        ref = P.propagate(ref_ei, ref, ref_w)
    assert_close(ops.spmm_k(Graph(ei, n, w), x.cuda(), 4), ref, 2e-5, "weighted chain")
