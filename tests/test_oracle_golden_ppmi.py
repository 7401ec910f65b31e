"""The oracle's PPMI restatement (oracle/ppmi.py, oracle/nn.py:PPMIConv) against vectors made by executing the
reference's own pygda/nn/ppmi_conv.py and UDAGCN(ppmi=True) with the same NumPy seed (tests/golden/ppmi.pt)."""
import numpy as np
import torch

from conftest import assert_close, load_golden
from oracle import ppmi as OP
from oracle.data import Data
from oracle.models import UDAGCN


def _sorted(ei, w, n):
    key = ei[0] * n + ei[1]
    order = torch.argsort(key, stable=True)
    return key[order], w[order]


def test_ppmi_norm_reproduces_the_reference():
    g = load_golden("ppmi")["norm"]
    for path_len, case in g["cases"].items():
        np.random.seed(case["np_seed"])
        ei, w = OP.ppmi_norm(g["edge_index"], g["num_nodes"], path_len)
        assert torch.equal(ei, case["edge_index_out"])                      # same walks, same dict order: bit-exact
        assert w.dtype == case["weight_out"].dtype == torch.float32
        assert_close(w, case["weight_out"], 1e-6, f"ppmi weights path_len={path_len}")


def test_counts_form_equals_the_dict_form():
    g = load_golden("ppmi")["norm"]
    np.random.seed(5)
    counters = OP.walk_counters(g["edge_index"], 5)
    ei, w = OP.ppmi_from_counters(counters, 5)
    a, b, c = [], [], []
    for s, ctr in counters.items():
        for t, n in ctr.items():
            a.append(s); b.append(t); c.append(n)
    w2 = OP.ppmi_from_counts(a, b, c, 5)
    assert np.allclose(w.numpy(), w2, rtol=1e-12, atol=1e-14)


def test_udagcn_ppmi_forward_model_reproduces_the_reference():
    g = load_golden("ppmi")["udagcn_ppmi"]
    est = UDAGCN(**g["hparams"])
    net = est.udagcn
    net.load_state_dict(g["state"])
    net.encoder.dropout_layers = [torch.nn.Identity() for _ in net.encoder.dropout_layers]
    net.ppmi_encoder.dropout_layers = [torch.nn.Identity() for _ in net.ppmi_encoder.dropout_layers]
    for m in net.models:
        m.eval()
    np.random.seed(g["np_seed"])
    src, tgt = Data(**g["source"]), Data(**g["target"])
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"], g["epoch"])
    for (li, name), (ei, w) in g["ppmi_caches"].items():
        got_ei, got_w = net.ppmi_encoder.conv_layers[li].cache_dict[name]
        assert torch.equal(got_ei, ei)
        assert_close(got_w, w, 1e-6, f"cached ppmi graph layer {li} {name}")
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-5, "target logits")
    net.zero_grad()
    loss.backward()
    for k, p in net.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)
