"""The paired bottleneck evaluations (ops.ActDropoutPairFn / PairGraphConvActFn, gda_spmm_nb_f32,
gda_bias_act_dropout_rep_*) against the one-at-a-time nodes they replace: the reference evaluates
feat_bottleneck twice per domain and step (pygda/models/a2gnn.py:181 & :192, :193 & :211); the
stacked form must give the values of two separate evaluations whose dropout masks are those of rows
[0, N) and [N, 2N) of the stacked matrix."""
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close

pytestmark = pytest.mark.gpu


def _graph(n, e, seed):
    from pygda_b200.graph import Graph
    from pygda_b200.synthetic import powerlaw_edge_index
    ei = powerlaw_edge_index(n, e, seed=seed, offset=2.0)
    return Graph(ei.cuda(), n), ei


@pytest.mark.parametrize("h", [128, 64, 5])
@pytest.mark.parametrize("transpose", [False, True])
def test_batched_aggregation_is_bit_identical_to_two_calls(h, transpose):
    """H = 128 takes the batched work-list kernel, the other widths the per-matrix fallback."""
    from pygda_b200 import ops
    n = 20000
    gr, _ = _graph(n, 400000, 2)                     # has hub rows (segments + ordered reduction)
    assert gr.num_long_rows > 0
    x = torch.randn(2 * n, h, device="cuda")
    y = ops.spmm(gr, x, transpose=transpose, nb=2)
    assert torch.equal(y[:n], ops.spmm(gr, x[:n].contiguous(), transpose=transpose))
    assert torch.equal(y[n:], ops.spmm(gr, x[n:].contiguous(), transpose=transpose))
    y2 = ops.spmm(gr, x, transpose=transpose, nb=2)
    assert torch.equal(y, y2)                        # counters reset, fixed summation order


@pytest.mark.parametrize("h", [128, 16])
def test_batched_epilogue_masks_are_those_of_the_stacked_matrix(h):
    from pygda_b200 import ops
    n = 5000
    gr, _ = _graph(n, 60000, 3)
    x = torch.randn(2 * n, h, device="cuda")
    b = torch.randn(h, device="cuda")
    seed = 0x1234567
    off = ops.dropout_rng.offset                     # device-side seed offset, if an earlier test enabled it
    y = ops.spmm(gr, x, bias=b, relu=True, dropout_p=0.5, seed=seed, nb=2, seed_offset=off)
    ya = ops.spmm(gr, x[:n].contiguous(), bias=b, relu=True, dropout_p=0.5, seed=seed, seed_offset=off)
    yb = ops.spmm(gr, x[n:].contiguous(), bias=b, relu=True, dropout_p=0.5, seed=ops.shift_seed(seed, n * h),
                  seed_offset=off)
    assert torch.equal(y[:n], ya) and torch.equal(y[n:], yb)
    assert not torch.equal((ya != 0), (yb != 0))     # independent masks
    # the stand-alone elementwise kernel draws the same mask for the same (seed, index)
    z = ops.spmm(gr, x, bias=b, relu=True, nb=2)
    zd = ops.ActDropoutFn.apply(z, 0, 0.5, seed)
    assert torch.equal(zd, y)


def test_k_steps_batched():
    from pygda_b200 import ops
    n = 4000
    gr, _ = _graph(n, 50000, 4)
    x = torch.randn(2 * n, 128, device="cuda")
    y = ops.spmm_k(gr, x, 10, nb=2)
    assert torch.equal(y[:n], ops.spmm_k(gr, x[:n].contiguous(), 10))
    assert torch.equal(y[n:], ops.spmm_k(gr, x[n:].contiguous(), 10))


@pytest.mark.parametrize("cols", [128, 6])
def test_act_dropout_pair_matches_two_single_nodes(cols):
    from pygda_b200 import ops
    n, seed, p = 3001, 987654321, 0.5
    x = torch.randn(n, cols, device="cuda", requires_grad=True)
    a, b = ops.ActDropoutPairFn.apply(x, 1, p, seed)
    x2 = x.detach().clone().requires_grad_(True)
    a2 = ops.ActDropoutFn.apply(x2, 1, p, seed)
    b2 = ops.ActDropoutFn.apply(x2, 1, p, ops.shift_seed(seed, n * cols))
    assert torch.equal(a, a2) and torch.equal(b, b2)
    ga, gb = torch.randn_like(a), torch.randn_like(b)
    (a * ga).sum().add((b * gb).sum()).backward()
    (a2 * ga).sum().add((b2 * gb).sum()).backward()
    assert_close(x.grad, x2.grad, 1e-6, "pair backward, both halves")
    # one half only: the other receives no gradient at all
    for use_a in (True, False):
        x3 = x.detach().clone().requires_grad_(True)
        a3, b3 = ops.ActDropoutPairFn.apply(x3, 1, p, seed)
        ((a3 * ga).sum() if use_a else (b3 * gb).sum()).backward()
        x4 = x.detach().clone().requires_grad_(True)
        t = ops.ActDropoutFn.apply(x4, 1, p, seed if use_a else ops.shift_seed(seed, n * cols))
        (t * (ga if use_a else gb)).sum().backward()
        assert torch.equal(x3.grad, x4.grad)


@pytest.mark.parametrize("k", [0, 3])
@pytest.mark.parametrize("use", ["both", "a", "b"])
def test_pair_conv_matches_two_convs(k, use):
    from pygda_b200 import ops
    n, fin, h, p, seed = 4000, 128, 128, 0.5, 424242
    gr, _ = _graph(n, 50000, 5)
    torch.manual_seed(0)
    w = (torch.randn(h, fin, device="cuda") * 0.1).requires_grad_(True)
    bias = torch.randn(h, device="cuda").requires_grad_(True)
    xa = torch.randn(n, fin, device="cuda", requires_grad=True)
    xb = torch.randn(n, fin, device="cuda", requires_grad=True)
    ya, yb = ops.PairGraphConvActFn.apply(xa, xb, w, bias, gr if k else None, k, False, 1, p, seed)
    ga, gb = torch.randn_like(ya), torch.randn_like(yb)
    loss = 0
    if use in ("both", "a"):
        loss = loss + (ya * ga).sum()
    if use in ("both", "b"):
        loss = loss + (yb * gb).sum()
    loss.backward()
    got = [t.grad.clone() if t.grad is not None else None for t in (xa, xb, w, bias)]

    w2, bias2 = w.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    xa2, xb2 = xa.detach().clone().requires_grad_(True), xb.detach().clone().requires_grad_(True)
    ra = ops.ActDropoutFn.apply(ops.graph_conv(xa2, w2, bias2, gr if k else None, k), 1, p, seed)
    rb = ops.ActDropoutFn.apply(ops.graph_conv(xb2, w2, bias2, gr if k else None, k), 1, p,
                                ops.shift_seed(seed, n * h))
    assert_close(ya, ra, 1e-6, "half a")
    assert_close(yb, rb, 1e-6, "half b")
    # same masks (a relu boundary may flip on a last-bit difference between the stacked and the single GEMM)
    assert int(((ya != 0) != (ra != 0)).sum()) <= 2 and int(((yb != 0) != (rb != 0)).sum()) <= 2
    loss = 0
    if use in ("both", "a"):
        loss = loss + (ra * ga).sum()
    if use in ("both", "b"):
        loss = loss + (rb * gb).sum()
    loss.backward()
    want = [t.grad for t in (xa2, xb2, w2, bias2)]
    for name, g, r in zip(("xa", "xb", "weight", "bias"), got, want):
        if r is None:
            assert g is None, f"{name}: gradient where none is due"
        else:
            assert_close(g, r, 2e-5, f"grad {name} ({use}, k={k})")


def test_bottleneck_pair_values_and_forward_model_step():
    """feat_bottleneck_pair == (feat_bottleneck, feat_bottleneck) without dropout; with dropout the
    estimator step runs through the paired path, gives finite results and learns."""
    from pygda_b200.models import A2GNN
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(3000, 30000, 200, 5, seed=3, target_nodes=2500, target_edges=24000, device="cuda:0")
    est = A2GNN(in_dim=200, hid_dim=128, num_classes=5, num_layers=3, dropout=0.5, s_pnums=0, t_pnums=4,
                weight=10, weight_decay=0.005, lr=0.01, epoch=200, device="cuda:0", verbose=0)
    torch.manual_seed(0)
    net = est.a2gnn = est.init_model()
    net.eval()
    f1, f2 = net.feat_bottleneck_pair(tgt.x, tgt.edge_index, None, 4)
    assert f1 is f2
    assert_close(f1, net.feat_bottleneck(tgt.x, tgt.edge_index, None, 4), 1e-6, "eval-mode pair")
    net.train()
    a, b = net.feat_bottleneck_pair(tgt.x, tgt.edge_index, None, 4)
    assert a.shape == b.shape == (2500, 128) and not torch.equal(a, b)
    keep = float((a != 0).float().mean())
    assert 0.15 < keep < 0.35                       # relu (~1/2) x dropout 0.5
    opt = Adam(net.parameters(), lr=0.01, weight_decay=0.005)
    losses = []
    for i in range(8):
        loss, s_logits, t_logits, _ = est.train_step(src, tgt, est.alpha_at(i, 200), opt)
        losses.append(loss.item())
        assert torch.isfinite(s_logits).all() and torch.isfinite(t_logits).all()
        assert t_logits.grad_fn is not None         # still part of the tape, like the reference's
    assert all(l == l and abs(l) < 1e6 for l in losses)
