"""Oracle AdaGCN / GNN against vectors produced by the reference's own files."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.data import Data
from oracle.models import AdaGCN as OracleAdaGCN, GNN as OracleGNN


@pytest.mark.parametrize("mode", ["node", "graph"])
def test_adagcn_forward_model(mode):
    g = load_golden("adagcn_" + mode)
    est = OracleAdaGCN(**g["hparams"])
    est.adagcn.load_state_dict(g["state"])
    est.discriminator.load_state_dict(g["critic_state"])
    est.adagcn.eval()
    est.discriminator.eval()
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(Data(**g["source"]), Data(**g["target"]))
    est.adagcn.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-5, "target logits")
    for k, v in est.discriminator.state_dict().items():
        assert_close(v, g["critic_state_after"][k], 1e-4, "critic after 10 iterations: " + k)
    for k, p in est.adagcn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)


def test_gnn_gcn_forward_model():
    g = load_golden("gnn_gcn")
    est = OracleGNN(**g["hparams"])
    est.gnn.load_state_dict(g["state"])
    est.gnn.train()
    loss, s_logits, t_logits = est.forward_model(Data(**g["source"]), Data(**g["target"]))
    est.gnn.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    for k, p in est.gnn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)
