"""Degenerate inputs through the newer entry points: empty edge lists, isolated nodes, all-zero feature
matrices, graphs without edges inside a mini-batch, single-row inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ppmi_of_a_graph_without_edges_is_loops_only():
    from pygda_b200.nn import PPMIConv
    from pygda_b200.ppmi import ppmi_edges, ppmi_walks
    n = 17
    empty = torch.zeros(2, 0, dtype=torch.int64, device="cuda")
    ei, w = ppmi_edges(empty, n, 5, seed=1)
    assert ei.shape == (2, 0) and w.numel() == 0
    assert bool((ppmi_walks(empty, n, 5, rounds=3, seed=1) == -1).all())
    conv = PPMIConv(4, 3).cuda()
    x = torch.randn(n, 4, device="cuda")
    y = conv(x, empty, "c")                              # norm() degenerates to the identity (remaining self loops)
    ref = x @ conv.weight + conv.bias
    assert torch.allclose(y, ref, atol=1e-5)


def test_ppmi_with_isolated_nodes_and_a_single_edge():
    from pygda_b200.ppmi import ppmi_edges
    ei = torch.tensor([[3], [8]], device="cuda")
    out, w, cnt = ppmi_edges(ei, 12, 4, rounds=40, seed=2, return_counts=True)
    pairs = set(zip(out[0].tolist(), out[1].tolist()))
    assert pairs == {(3, 8), (3, 3), (8, 3), (8, 8)}      # the walk bounces between the two endpoints
    assert int(cnt.sum()) >= 2 * 40 and bool((w >= 0).all())


def test_batched_aggregation_on_a_loops_only_graph():
    from pygda_b200 import ops
    from pygda_b200.graph import Graph
    n = 1000
    g = Graph(torch.zeros(2, 0, dtype=torch.int64, device="cuda"), n)      # A_hat = I
    x = torch.randn(2 * n, 128, device="cuda")
    assert torch.equal(ops.spmm(g, x, nb=2), x)
    assert torch.equal(ops.spmm_k(g, x, 5, nb=2), x)


def test_packed_staging_of_an_all_zero_matrix_and_a_single_row():
    from pygda_b200.data import Data
    for x in (torch.zeros(50, 300), torch.tensor([[0.0, 2.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]])):
        d = Data(x=x, edge_index=torch.zeros(2, 0, dtype=torch.long), y=torch.zeros(x.size(0), dtype=torch.long))
        p = d.pin_memory(pack=True)
        assert "_packed_x" in p.__dict__
        assert torch.equal(p.to("cuda:0").x.cpu(), x)


def test_collating_graphs_without_edges():
    from pygda_b200.data import Batch, Data, DeviceGraphDataset
    graphs = [Data(x=torch.randn(3, 5), edge_index=torch.tensor([[0, 1], [1, 0]]), y=torch.tensor([1])),
              Data(x=torch.randn(1, 5), edge_index=torch.zeros(2, 0, dtype=torch.long), y=torch.tensor([0])),
              Data(x=torch.randn(4, 5), edge_index=torch.tensor([[0, 3, 2], [3, 0, 2]]), y=torch.tensor([1]))]
    res = DeviceGraphDataset(graphs, "cuda:0")
    for ids in ([1], [1, 1], [2, 1, 0]):
        got, ref = res.collate(ids), Batch.from_data_list([graphs[i] for i in ids])
        assert torch.equal(got.x.cpu(), ref.x) and torch.equal(got.edge_index.cpu(), ref.edge_index)
        assert torch.equal(got.batch.cpu(), ref.batch) and torch.equal(got.y.cpu(), ref.y)


def test_confusion_of_zero_rows_and_attention_of_one_row():
    from pygda_b200.metrics import confusion_from_logits, f1_from_confusion
    from pygda_b200.nn import Attention
    cm = confusion_from_logits(torch.zeros(0, dtype=torch.long).cuda(), torch.zeros(0, 4).cuda())
    assert cm.shape == (4, 4) and int(cm.sum()) == 0 and f1_from_confusion(cm, "micro") == 0.0
    att = Attention(8).cuda()
    a, b = torch.randn(1, 8, device="cuda"), torch.randn(1, 8, device="cuda")
    out = att([a, b])
    s = torch.stack([a, b], 1)
    ref = (s * torch.softmax(att.dense_weight(s), 1)).sum(1)
    assert torch.allclose(out, ref, atol=1e-6)


def test_bernprop_on_a_graph_without_edges():
    from pygda_b200.nn import BernProp
    prop = BernProp(3).cuda()
    x = torch.randn(9, 4, device="cuda")
    y = prop(x, torch.zeros(2, 0, dtype=torch.int64, device="cuda"))
    # L = I, 2I - L = I: sum_k C(3,k)/8 * 1 * x = x
    assert torch.allclose(y, x, atol=1e-6)
