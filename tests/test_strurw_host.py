"""Host-side logic of the StruRW path that needs no GPU: the edge re-weighting (pure integer / float64 index work,
bit-exact against the vectors made by the reference's own cal_reweight), the folded aggregation weights of the
re-weighted layers (``message_values``) against the reference layers' outputs, the mixup node relabelling, and the
loader semantics the re-weighting depends on."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import pyg_ops as P


def _data(d, **kw):
    from pygda_b200.data import Data
    return Data(**d, **kw)


def test_cal_reweight_bit_exact_on_the_host():
    from pygda_b200.models import StruRW
    g = load_golden("strurw")["reweight"]
    est = StruRW(in_dim=12, hid_dim=8, num_classes=g["num_classes"], device="cpu")
    s, t = _data(g["source"]), _data(g["target"])
    est.cal_reweight(s, t, g["target_pred"])
    assert s.edge_weight.dtype == torch.float32
    assert torch.equal(s.edge_weight, g["edge_weight"])


def test_edge_probabilities_equal_the_dense_adjacency_form():
    """cal_edge_prob_sep restated with the reference's dense N x N adjacency (strurw.py:509-548), float64."""
    from pygda_b200.models import StruRW
    g = load_golden("strurw")["reweight"]
    C = g["num_classes"]
    est = StruRW(in_dim=12, hid_dim=8, num_classes=C, device="cpu")
    s, t = _data(g["source"]), _data(g["target"])
    got = est.cal_edge_prob_sep(s, t, g["target_pred"])

    def dense(d, labels):
        n = d.x.shape[0]
        adj = torch.zeros(n, n, dtype=torch.float64)
        adj.index_put_((d.edge_index[0], d.edge_index[1]), torch.ones(d.edge_index.size(1), dtype=torch.float64),
                       accumulate=True)
        onehot = torch.nn.functional.one_hot(labels, C).double()
        cnt = onehot.sum(0)
        return onehot.t() @ adj @ onehot, torch.outer(cnt, cnt)

    num, den = dense(s, s.y)
    assert torch.equal(torch.nan_to_num(got[0], nan=-1.0), torch.nan_to_num(num / den, nan=-1.0))
    num, den = dense(t, g["target_pred"])
    assert torch.equal(got[1], num / (den + 1e-12))
    num, den = dense(t, t.y)
    assert torch.equal(torch.nan_to_num(got[2], nan=-1.0), torch.nan_to_num(num / den, nan=-1.0))


@pytest.mark.parametrize("name", ["gcn_mean", "gcn_add"])
def test_folded_weights_reproduce_the_reference_gcn_layers(name):
    """out = A (x W^T) + b with A built from ``message_values`` equals the reference layer's output."""
    from pygda_b200.nn.reweight_gnn import message_values
    g = load_golden("strurw")["layers"]
    c = g["cases"][name]
    ei, n = g["edge_index"], g["num_nodes"]
    aggr = name.split("_")[1]
    w_norm = P.gcn_norm_by_col(ei, None, n, False, False)[1] if aggr == "mean" else None
    val = message_values(ei, w_norm, g["edge_weight"], g["lmda"], n, aggr, to_source=True)
    h = g["x"] @ c["state"]["lin.weight"].t()
    out = P.scatter_add(val.view(-1, 1) * h[ei[1]], ei[0], 0, n) + c["state"]["bias"]
    assert_close(out, c["y"], 1e-6, name)


@pytest.mark.parametrize("name", ["gs_mean", "gs_add"])
def test_folded_weights_reproduce_the_reference_gs_layers(name):
    """Linear-per-node instead of per-edge, split agg_lin weight instead of the concatenation."""
    from pygda_b200.nn.reweight_gnn import message_values
    g = load_golden("strurw")["layers"]
    c = g["cases"][name]
    st = c["state"]
    ei, n = g["edge_index"], g["num_nodes"]
    val = message_values(ei, None, g["edge_weight"], g["lmda"], n, name.split("_")[1], to_source=True)
    h = g["x"] @ st["lin.weight"].t() + st["lin.bias"]
    a = P.scatter_add(val.view(-1, 1) * h[ei[1]], ei[0], 0, n)
    w = st["agg_lin.weight"]
    out = torch.relu(a @ w[:, :7].t() + st["agg_lin.bias"] + g["x"] @ w[:, 7:].t())
    assert_close(out, c["y"], 1e-6, name)


def test_folded_weights_reproduce_the_reference_mixup_conv():
    from pygda_b200.nn.reweight_gnn import message_values
    g = load_golden("strurw")["layers"]
    c = g["cases"]["mixup_conv"]
    st = c["state"]
    ei, n = g["edge_index"], g["num_nodes"]
    w_norm = P.gcn_norm_by_col(ei, None, n, False, False)[1]
    val = message_values(ei, w_norm, g["edge_weight"], g["lmda"], n, "add", to_source=False)
    h = g["x"] @ st["lin.weight"].t()
    out = P.scatter_add(val.view(-1, 1) * h[ei[0]], ei[1], 0, n) + c["x_cen"] @ st["lin_cen.weight"].t() + st["bias"]
    assert_close(out, c["y"], 1e-6, "MixUpGCNConv")


def test_shuffle_data_consumes_np_random_like_the_reference_and_relabels_edges():
    from oracle.models import StruRW as OStruRW
    from pygda_b200.models import StruRW
    g = load_golden("strurw")["reweight"]
    s = _data(g["source"])
    est = StruRW(in_dim=12, hid_dim=8, num_classes=4, device="cpu")
    np.random.seed(11)
    data_b, perm = est.shuffle_data(s)
    after = np.random.rand()
    np.random.seed(11)
    ei_ref, perm_ref = OStruRW.shuffle_edges(s.edge_index, s.x.shape[0])
    assert np.array_equal(perm, perm_ref) and after == np.random.rand()
    assert torch.equal(data_b.edge_index, ei_ref)
    assert data_b.x is None and torch.equal(data_b.y, s.y[torch.from_numpy(perm)])
    # an edge (u, v) of the shuffled graph is the edge (perm[u], perm[v]) of the original one
    p = torch.from_numpy(perm)
    assert torch.equal(torch.stack([p[data_b.edge_index[0]], p[data_b.edge_index[1]]]), s.edge_index)


def test_loader_permutes_edge_weights_with_their_edges_and_yields_fresh_batches():
    from pygda_b200.data import Data, NeighborLoader
    ei = torch.tensor([[0, 1, 2, 3, 1], [2, 0, 1, 0, 3]])
    w = torch.tensor([10., 11., 12., 13., 14.])
    d = Data(x=torch.randn(4, 3), edge_index=ei, y=torch.zeros(4, dtype=torch.long), edge_weight=w)
    loader = NeighborLoader(d, [-1, -1], batch_size=4)
    b1 = next(iter(loader))
    order = torch.argsort(ei[1], stable=True)
    assert torch.equal(b1.edge_index, ei[:, order]) and torch.equal(b1.edge_weight, w[order])
    b1.edge_weight = torch.zeros(5)                    # what StruRW.cal_reweight does to the batch it was handed
    b2 = next(iter(loader))
    assert b2 is not b1 and torch.equal(b2.edge_weight, w[order])      # PyG: a new Data per batch
    assert b2.edge_index is b1.edge_index              # same tensors: graph caches keyed by identity keep hitting


def test_constructor_attributes_and_mode_check():
    from pygda_b200.models import StruRW
    est = StruRW(in_dim=5, hid_dim=4, num_classes=3, device="cpu")
    assert (est.gnn, est.pooling, est.mode, est.lamb, est.ew_start, est.ew_freq) == ("GS", "mean", "erm", 0.8, 100, 20)
    assert est.reweight is True and est.pseudo is True and est.cls_dim == 128 and est.cls_layers == 2
    assert est.lr == 0.05 and est.weight_decay == 0.0001 and est.epoch == 100
    with pytest.raises(AssertionError, match="unsupport training mode"):
        StruRW(in_dim=5, hid_dim=4, num_classes=3, mode="nope", device="cpu")


def test_reweight_schedule():
    """(epoch + 1) >= ew_start and (epoch + 1) % ew_freq == 0 with pseudo labels; once at ew_start - 1 without."""
    from pygda_b200.models import StruRW
    fired = []

    class Probe(StruRW):
        def cal_reweight(self, s, t, p):
            fired.append(self._epoch)

    for pseudo, expect in ((True, [5, 8]), (False, [3])):
        fired.clear()
        est = Probe(in_dim=5, hid_dim=4, num_classes=3, ew_start=4, ew_freq=3, pseudo=pseudo, device="cpu")
        for epoch in range(10):
            est._epoch = epoch
            est._maybe_reweight(None, None, None, epoch)
        assert fired == expect


def test_structure_difference_diagnostics_equal_the_reference_formulas():
    """cal_str_dif_rel / cal_str_diff_ratio (strurw.py:550-614) restated here op by op, incl. zeros in the tables."""
    from pygda_b200.models import StruRW
    est = StruRW(in_dim=5, hid_dim=4, num_classes=3, device="cpu")
    torch.manual_seed(3)
    pred, true = torch.rand(4, 4, dtype=torch.float64), torch.rand(4, 4, dtype=torch.float64)
    pred[0, 1] = 0.0
    true[2, 3] = 0.0
    true[1, 1] = 0.0
    pred[3, 3] = 0.0
    pred[1, 2] = true[1, 2] = 0.0

    cls1, cls0 = (pred - true).abs(), ((1 - pred) - (1 - true)).abs()
    abs_diff = 0.5 * cls0 + 0.5 * cls1
    r1, r2 = abs_diff / true, abs_diff / pred
    rel = 0.5 * r1 + 0.5 * r2
    rel[torch.isinf(r1)] = r2[torch.isinf(r1)]
    rel[torch.isinf(r2)] = r1[torch.isinf(r2)]
    rel[torch.isnan(rel)] = 0
    got = est.cal_str_dif_rel(pred, true)
    assert torch.equal(got[0], abs_diff.sum() / 16) and torch.equal(got[1], rel.sum() / 16)

    def intra(m):
        d = torch.diagonal(m, 0).repeat_interleave(m.size(1)).view(-1, m.size(1))
        r = torch.div(m, d)
        r[torch.isnan(r)] = 1
        r[torch.isinf(r)] = m[torch.isinf(r)]
        return r
    ratio = torch.div(intra(pred), intra(true))
    ratio[torch.isnan(ratio)] = 1
    ratio[torch.isinf(ratio)] = 1
    want = (ratio.sum() - torch.diagonal(ratio).sum()) / (16 - 4)
    assert torch.equal(est.cal_str_diff_ratio(pred, true), want)


def test_reweighting_degenerate_inputs():
    """No edges at all; a class that occurs in neither graph; a class the target never predicts."""
    from pygda_b200.data import Data
    from pygda_b200.models import StruRW
    est = StruRW(in_dim=3, hid_dim=4, num_classes=4, device="cpu")
    x = torch.zeros(6, 3)
    empty = torch.zeros(2, 0, dtype=torch.long)
    s = Data(x=x, edge_index=empty, y=torch.tensor([0, 0, 1, 1, 2, 2]))
    t = Data(x=x, edge_index=empty, y=torch.tensor([0, 1, 2, 0, 1, 2]))
    est.cal_reweight(s, t, torch.tensor([0, 0, 0, 1, 1, 1]))
    assert s.edge_weight.shape == (0,) and s.edge_weight.dtype == torch.float32
    ei = torch.tensor([[0, 2, 4, 1, 0], [2, 4, 0, 3, 1]])
    s = Data(x=x, edge_index=ei, y=torch.tensor([0, 0, 1, 1, 2, 2]))
    t = Data(x=x, edge_index=ei, y=torch.tensor([0, 1, 2, 0, 1, 2]))
    src_p, tgt_p, true_p = est.cal_edge_prob_sep(s, t, torch.tensor([0, 0, 1, 1, 1, 1]))
    assert torch.isnan(src_p[3]).all() and torch.isnan(src_p[:, 3]).all()       # class 3 absent: 0 / 0
    assert (tgt_p[2] == 0).all() and (tgt_p[3] == 0).all()                       # never predicted: 0 / 1e-12
    est.cal_reweight(s, t, torch.tensor([0, 0, 1, 1, 1, 1]))
    assert torch.isfinite(s.edge_weight).all() and s.edge_weight.shape == (5,)
    # every edge weight is the table entry of its (class of edge_index[1], class of edge_index[0]) pair
    R = tgt_p / src_p
    R[torch.isinf(R)] = 1
    R[torch.isnan(R)] = 1
    assert torch.equal(s.edge_weight, R[s.y[ei[1]], s.y[ei[0]]].float())


def test_message_values_degenerate_inputs():
    from pygda_b200.nn.reweight_gnn import message_values
    empty = torch.zeros(2, 0, dtype=torch.long)
    assert message_values(empty, None, torch.zeros(0), 0.5, 4, "mean").shape == (0,)
    ei = torch.tensor([[0, 0, 0, 2], [1, 1, 3, 2]])                               # duplicate edge, self loop, isolated node 1/3 rows
    v = message_values(ei, None, torch.tensor([1., 3., 1., 2.]), 0.5, 4, "mean")
    assert torch.allclose(v, torch.tensor([1 / 3, 2 / 3, 1 / 3, 1.5]))
    v = message_values(ei, None, None, 0.5, 4, "add")
    assert torch.equal(v, torch.ones(4))
    with pytest.raises(ValueError, match="unsupported aggregation"):
        message_values(ei, None, None, 0.5, 4, "max")


def test_fill_empty_rows_keeps_the_matrix_and_leaves_no_row_or_column_empty():
    from pygda_b200.nn.reweight_gnn import fill_empty_rows
    n = 7
    ei = torch.tensor([[0, 0, 2, 4], [1, 2, 0, 4]])           # 3, 5, 6 isolated; 1 has no out-edge; 4 only a loop
    val = torch.tensor([0.5, 1.5, 2.0, 3.0])
    ei2, val2 = fill_empty_rows(ei, val, n)
    dense = torch.zeros(n, n).index_put_((ei[0], ei[1]), val, accumulate=True)
    dense2 = torch.zeros(n, n).index_put_((ei2[0], ei2[1]), val2, accumulate=True)
    assert torch.equal(dense, dense2)                          # zero-weight entries only
    assert torch.equal(ei2[:, :4], ei) and torch.equal(val2[:4], val)
    for r in (0, 1):
        assert torch.bincount(ei2[r], minlength=n).min() >= 1
    assert sorted(ei2[0, 4:].tolist()) == [1, 3, 5, 6] and torch.equal(ei2[0, 4:], ei2[1, 4:])
    same = fill_empty_rows(ei2, val2, n)
    assert same[0] is ei2 and same[1] is val2                  # nothing to add the second time
