"""Graph-level mode of A2GNN / UDAGCN / GRADE on the GPU (segment-mean pooling kernels, per-graph labels and MMD rows)
against vectors made by executing the reference's own files on two collated graph batches
(tests/golden/graph_mode.pt)."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _check(net, g, loss, s_logits, t_logits):
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    net.zero_grad()
    loss.backward()
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    for k, v in got.items():
        assert_close(v, g["grads"][k], 2e-4, "grad " + k)


def _batches(G):
    from pygda_b200.data import Data
    return Data(**G["source"]).to("cuda:0"), Data(**G["target"]).to("cuda:0")


def test_a2gnn_graph_mode_golden():
    from pygda_b200.models import A2GNN
    G = load_golden("graph_mode")
    g = G["a2gnn"]
    est = A2GNN(device="cuda:0", verbose=0, **g["hparams"])
    est.a2gnn = est.init_model()
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.train()
    src, tgt = _batches(G)
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"])
    _check(est.a2gnn, g, loss, s_logits, t_logits)


def test_a2gnn_adversarial_graph_mode_raises_like_the_reference():
    from pygda_b200.models import A2GNN
    G = load_golden("graph_mode")
    e = G["a2gnn_adv_error"]
    est = A2GNN(device="cuda:0", verbose=0, **e["hparams"])
    est.a2gnn = est.init_model()
    src, tgt = _batches(G)
    with pytest.raises(ValueError) as info:
        est.forward_model(src, tgt, 0.2)
    assert str(info.value) == e["message"]


def test_udagcn_graph_mode_golden():
    from pygda_b200.models import UDAGCN
    G = load_golden("graph_mode")
    g = G["udagcn"]
    est = UDAGCN(device="cuda:0", verbose=0, **g["hparams"])
    est.udagcn = est.init_model()
    est.udagcn.load_state_dict(g["state"])
    est.udagcn.encoder.dropout_p = [0.0 for _ in est.udagcn.encoder.dropout_p]   # see make_golden_graph_mode.py
    est._set_train(False)
    src, tgt = _batches(G)
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"], g["epoch"])
    _check(est.udagcn, g, loss, s_logits, t_logits)


@pytest.mark.parametrize("disc", ["js", "mmd"])
def test_grade_graph_mode_golden(disc):
    from pygda_b200.models import GRADE
    G = load_golden("graph_mode")
    g = G["grade_" + disc]
    est = GRADE(device="cuda:0", verbose=0, **g["hparams"])
    est.grade = est.init_model()
    est.grade.load_state_dict(g["state"])
    est.grade.train()
    src, tgt = _batches(G)
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"])
    _check(est.grade, g, loss, s_logits, t_logits)
