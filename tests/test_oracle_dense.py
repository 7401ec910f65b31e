"""Pin the oracle's restatement of the upstream (PyG / torch_scatter) ops against
dense fp64 linear algebra (oracle plan item 1)."""
import torch

from oracle import pyg_ops as P
from oracle import nn as ONN
from oracle import mmd as OM
from conftest import assert_close


def rand_graph(n, e, seed=0, loops=True):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(n, (2, e), generator=g)
    if not loops:
        ei = ei[:, ei[0] != ei[1]]
    return ei


def test_add_remaining_self_loops_order_and_weights():
    ei = torch.tensor([[0, 1, 2, 2, 3, 1], [1, 0, 2, 3, 2, 1]])
    w = torch.tensor([1., 2., 7., 3., 4., 9.])
    out_ei, out_w = P.add_remaining_self_loops(ei, w, 1.0, 5)
    assert out_ei.tolist() == [[0, 1, 2, 3, 0, 1, 2, 3, 4], [1, 0, 3, 2, 0, 1, 2, 3, 4]]
    assert out_w.tolist() == [1., 2., 3., 4., 1., 9., 7., 1., 1.]


def test_propagate_equals_dense_matmul():
    n, h = 50, 7
    ei = rand_graph(n, 300, 1)
    w = torch.rand(ei.size(1), dtype=torch.float64)
    x = torch.randn(n, h, dtype=torch.float64)
    a = P.dense_adj(ei, w, n)
    assert_close(P.propagate(ei, x, w), a @ x, 1e-12, "propagate")


def test_gcn_norm_by_col_is_symmetric_normalisation():
    n = 40
    ei = rand_graph(n, 200, 2, loops=False)
    ei = torch.cat([ei, ei.flip(0)], 1)
    ei2, w = P.gcn_norm_by_col(ei, None, n, dtype=torch.float64)
    a = P.dense_adj(ei, torch.ones(ei.size(1)), n) + torch.eye(n, dtype=torch.float64)
    d = a.sum(1)
    ref = a / d.sqrt().view(-1, 1) / d.sqrt().view(1, -1)
    assert_close(P.dense_adj(ei2, w, n), ref, 1e-12, "A_hat")


def test_gcn_norm_by_row_vs_by_col_differ_only_for_directed():
    n = 30
    ei = rand_graph(n, 150, 3, loops=False)
    _, wc = P.gcn_norm_by_col(ei, None, n)
    _, wr = P.gcn_norm_by_row(ei, n)
    sym = torch.cat([ei, ei.flip(0)], 1)
    _, wc2 = P.gcn_norm_by_col(sym, None, n)
    _, wr2 = P.gcn_norm_by_row(sym, n)
    assert not torch.allclose(wc, wr)
    assert torch.allclose(wc2, wr2)


def test_prop_gcn_conv_is_Ak_XW_plus_b():
    torch.manual_seed(0)
    n, f, h = 35, 9, 5
    ei = rand_graph(n, 160, 4)
    conv = ONN.PropGCNConv(f, h).double()
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    x = torch.randn(n, f, dtype=torch.float64)
    ei2, w = P.gcn_norm_by_col(ei, None, n, dtype=torch.float64)
    a = P.dense_adj(ei2, w, n)
    for k in (0, 1, 4):
        ref = torch.linalg.matrix_power(a, k) @ (x @ conv.lin.weight.t()) + conv.bias
        assert_close(conv(x, ei, k), ref, 1e-12, f"k={k}")


def test_global_mean_pool():
    x = torch.arange(12.).view(6, 2)
    b = torch.tensor([0, 0, 1, 1, 1, 3])
    out = P.global_mean_pool(x, b)
    assert out.shape == (4, 2)
    assert torch.allclose(out[0], x[:2].mean(0)) and torch.allclose(out[1], x[2:5].mean(0))
    assert torch.all(out[2] == 0) and torch.allclose(out[3], x[5])


def test_mmd_blocked_equals_broadcast_and_closed_form_bandwidth():
    torch.manual_seed(1)
    s, t = torch.randn(40, 6, dtype=torch.float64), torch.randn(40, 6, dtype=torch.float64) + 0.5
    a = OM.get_mmd(s, t, sqdist=OM.pairwise_sqdist_broadcast)
    b = OM.get_mmd(s, t, sqdist=lambda z: OM.pairwise_sqdist_blocked(z, 16))
    assert_close(a, b, 1e-12, "blocked L2")
    total = torch.cat([s, t])
    n = total.size(0)
    closed = 2 * n * (total ** 2).sum() - 2 * (total.sum(0) ** 2).sum()
    assert_close(OM.pairwise_sqdist_broadcast(total).sum(), closed, 1e-10, "sum L2 closed form")


def test_grad_reverse():
    x = torch.randn(4, 3, requires_grad=True)
    ONN.GradReverse.apply(x, 0.25).sum().backward()
    assert torch.allclose(x.grad, torch.full_like(x, -0.25))


def test_f1_from_confusion_equals_sklearn():
    """Host half of the device-side training score (pygda_b200/metrics): sklearn's micro / macro F1 from counts."""
    from sklearn.metrics import confusion_matrix, f1_score
    from pygda_b200.metrics import f1_from_confusion
    g = torch.Generator().manual_seed(0)
    for c, present in ((2, 2), (5, 5), (7, 4)):              # (7, 4): three classes absent from labels AND predictions
        y = torch.randint(present, (500,), generator=g)
        p = torch.randint(present, (500,), generator=g)
        cm = torch.from_numpy(confusion_matrix(y.numpy(), p.numpy(), labels=list(range(c))))
        for avg in ("micro", "macro"):
            assert abs(f1_from_confusion(cm, avg) - f1_score(y.numpy(), p.numpy(), average=avg)) < 1e-12
