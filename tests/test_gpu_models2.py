"""AdaGCN / GNN estimators on the GPU vs vectors from the reference's own forward_model."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["node", "graph"])
def test_adagcn_forward_model_golden(mode):
    from pygda_b200.data import Data
    from pygda_b200.models import AdaGCN
    g = load_golden("adagcn_" + mode)
    est = AdaGCN(device="cuda:0", verbose=0, **g["hparams"])
    est.adagcn = est.init_model()
    est.adagcn.load_state_dict(g["state"])
    est.init_critic()
    est.discriminator.load_state_dict(g["critic_state"])
    est.adagcn.eval()
    est.discriminator.eval()
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    torch.manual_seed(g["seed"])          # gradient_penalty's torch.rand draws come from the CPU generator
    loss, s_logits, t_logits = est.forward_model(src, tgt)
    est.adagcn.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    for k, v in est.discriminator.state_dict().items():
        assert_close(v, g["critic_state_after"][k], 2e-4, "critic after 10 iterations: " + k)
    for k, p in est.adagcn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 2e-4, "grad " + k)


def test_gnn_gcn_forward_model_golden():
    from pygda_b200.data import Data
    from pygda_b200.models import GNN
    g = load_golden("gnn_gcn")
    est = GNN(device="cuda:0", verbose=0, **g["hparams"])
    est.gnn = est.init_model()
    est.gnn.load_state_dict(g["state"])
    est.gnn.train()
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    loss, s_logits, t_logits = est.forward_model(src, tgt)
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    for k, p in est.gnn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)


def test_adagcn_graph_level_fit_predict():
    from pygda_b200.models import AdaGCN
    from pygda_b200.synthetic import graph_dataset
    src = graph_dataset(96, 30, 2.05, 14, 2, seed=1)
    tgt = graph_dataset(64, 39, 3.7, 14, 2, seed=2)
    torch.manual_seed(0)
    model = AdaGCN(in_dim=14, hid_dim=32, num_classes=2, mode="graph", num_layers=2, epoch=2, batch_size=32,
                   lr=0.01, weight_decay=0.01, domain_weight=0.1, gp_weight=5, device="cuda:0", verbose=0)
    model.fit(src, tgt)
    logits, labels = model.predict(None)
    assert logits.shape == (64, 2) and labels.shape == (64,)
    # gnn_type='ppmi' (adagcn_base.py:53-57): PPMIConv layers, node level
    from pygda_b200.nn import PPMIConv
    from pygda_b200.synthetic import domain_pair
    s2, t2 = domain_pair(600, 4000, 14, 2, seed=3)
    pm = AdaGCN(in_dim=14, hid_dim=8, num_classes=2, num_layers=2, gnn_type="ppmi", epoch=1, device="cuda:0", verbose=0)
    pm.fit(s2, t2)
    assert all(isinstance(c, PPMIConv) for c in pm.adagcn.encoder.conv_layers)
    out, _ = pm.predict(t2)
    assert out.shape == (600, 2) and torch.isfinite(out).all()


def test_gnn_fit_predict_uses_data():
    from pygda_b200.models import GNN
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1200, 9000, 32, 3, seed=2, device="cuda:0")
    model = GNN(in_dim=32, hid_dim=16, num_classes=3, num_layers=2, epoch=3, device="cuda:0", verbose=0)
    model.fit(src, tgt)
    logits, labels = model.predict(tgt)
    assert logits.shape == (1200, 3) and torch.equal(labels, tgt.y)
    assert torch.allclose(logits.exp().sum(1), torch.ones(1200, device="cuda:0"), atol=1e-4)   # log_softmax output
