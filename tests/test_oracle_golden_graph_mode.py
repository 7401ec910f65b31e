"""Graph-level mode of the A2GNN / UDAGCN / GRADE restatements (oracle/models.py, oracle/nn.py: pooled encodings,
per-graph labels and MMD rows) against vectors made by executing the reference's own files on two collated graph
batches (tests/golden/graph_mode.pt, tests/golden/make_golden_graph_mode.py)."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.data import Data
from oracle.models import A2GNN, GRADE, UDAGCN


def _check(net, g, loss, s_logits, t_logits, tol=1e-5):
    assert_close(loss, g["loss"], tol, "loss")
    assert_close(s_logits, g["source_logits"], tol, "source logits")
    assert_close(t_logits, g["target_logits"], tol, "target logits")
    net.zero_grad()
    loss.backward()
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    for k, v in got.items():
        assert_close(v, g["grads"][k], 1e-4, "grad " + k)


def test_a2gnn_graph_mode():
    G = load_golden("graph_mode")
    g = G["a2gnn"]
    est = A2GNN(**g["hparams"])
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.train()
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(Data(**G["source"]), Data(**G["target"]), g["alpha"])
    assert s_logits.shape[0] == G["source"]["num_graphs"] and t_logits.shape[0] == G["target"]["num_graphs"]
    _check(est.a2gnn, g, loss, s_logits, t_logits)


def test_a2gnn_adversarial_graph_mode_raises_like_the_reference():
    G = load_golden("graph_mode")
    e = G["a2gnn_adv_error"]
    assert e is not None and e["type"] == "ValueError"
    est = A2GNN(**e["hparams"])
    with pytest.raises(ValueError) as info:
        est.forward_model(Data(**G["source"]), Data(**G["target"]), 0.2)
    assert str(info.value) == e["message"]


def test_udagcn_graph_mode():
    G = load_golden("graph_mode")
    g = G["udagcn"]
    est = UDAGCN(**g["hparams"])
    est.udagcn.load_state_dict(g["state"])
    est.udagcn.encoder.dropout_layers = [torch.nn.Identity() for _ in est.udagcn.encoder.dropout_layers]
    for m in est.udagcn.models:
        m.eval()
    loss, s_logits, t_logits = est.forward_model(Data(**G["source"]), Data(**G["target"]), g["alpha"], g["epoch"])
    _check(est.udagcn, g, loss, s_logits, t_logits)


@pytest.mark.parametrize("disc", ["js", "mmd"])
def test_grade_graph_mode(disc):
    G = load_golden("graph_mode")
    g = G["grade_" + disc]
    est = GRADE(**g["hparams"])
    est.grade.load_state_dict(g["state"])
    est.grade.train()
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(Data(**G["source"]), Data(**G["target"]), g["alpha"])
    _check(est.grade, g, loss, s_logits, t_logits)
