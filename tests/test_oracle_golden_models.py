"""Oracle UDAGCN / GRADE against vectors produced by the reference's own files."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.data import Data
from oracle.models import GRADE as OracleGRADE, UDAGCN as OracleUDAGCN


def test_udagcn_forward_model():
    g = load_golden("udagcn")
    est = OracleUDAGCN(**g["hparams"])
    est.udagcn.load_state_dict(g["state"])
    est.udagcn.encoder.dropout_layers = [torch.nn.Identity() for _ in est.udagcn.encoder.dropout_layers]
    for m in est.udagcn.models:
        m.eval()
    loss, s_logits, t_logits = est.forward_model(Data(**g["source"]), Data(**g["target"]), g["alpha"], g["epoch"])
    est.udagcn.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-5, "target logits")
    for k, p in est.udagcn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)


@pytest.mark.parametrize("disc", ["js", "mmd", "c"])
def test_grade_forward_model(disc):
    g = load_golden("grade_" + disc)
    est = OracleGRADE(**g["hparams"])
    est.grade.load_state_dict(g["state"])
    est.grade.train()
    est.mmd_indices = (g["source_idx"], g["target_idx"])
    loss, s_logits, t_logits = est.forward_model(Data(**g["source"]), Data(**g["target"]), g["alpha"])
    est.grade.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-5, "target logits")
    for k, p in est.grade.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)
