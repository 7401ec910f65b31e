"""DGSDA's Bernstein propagation on the GPU (pygda_b200/nn/dgsda_base.py: two Horner sweeps, 2K aggregation
launches) against vectors made by executing the reference's own pygda/nn/dgsda_base.py / pygda/models/dgsda.py
(K + K(K+1)/2 propagations): same polynomial, different evaluation order -- fp32 tolerance 1e-5 / 1e-4."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [1, 4, 8])
def test_bernprop_matches_the_reference_vectors(K):
    from pygda_b200.nn import BernProp
    g = load_golden("dgsda")["bernprop"]
    c = g["cases"][K]
    prop = BernProp(K).cuda()
    with torch.no_grad():
        prop.temp.copy_(c["temp"])
    x = c["x"].cuda().requires_grad_(True)
    y = prop(x, g["edge_index"].cuda())
    y.backward(c["gout"].cuda())
    assert_close(y, c["y"], 1e-5, f"BernProp K={K}")
    assert_close(x.grad, c["gx"], 1e-5, "input gradient")
    assert_close(prop.temp.grad, c["gtemp"], 1e-4, "temp gradient")
    assert float(prop.temp.grad[-1]) == 0.0                  # temp_K < 0: relu gate closed


def test_bernprop_at_scale_against_the_oracle_evaluation_order():
    from oracle import nn as ONN
    from pygda_b200.nn import BernProp
    from pygda_b200.synthetic import powerlaw_edge_index
    n, h, K = 5000, 128, 6
    ei = powerlaw_edge_index(n, 50000, seed=4, offset=2.0)
    torch.manual_seed(0)
    x = torch.randn(n, h)
    temp = torch.rand(K + 1) + 0.1
    ora = ONN.BernProp(K)
    with torch.no_grad():
        ora.temp.copy_(temp)
    xr = x.clone().requires_grad_(True)
    yr = ora(xr, ei)
    go = torch.randn(n, h)
    yr.backward(go)
    prop = BernProp(K).cuda()
    with torch.no_grad():
        prop.temp.copy_(temp)
    xg = x.cuda().requires_grad_(True)
    yg = prop(xg, ei.cuda())
    yg.backward(go.cuda())
    assert_close(yg, yr, 2e-5, "forward")
    assert_close(xg.grad, xr.grad, 2e-5, "input gradient")
    assert_close(prop.temp.grad, ora.temp.grad, 1e-4, "temp gradient")


def test_dgsda_forward_model_golden():
    from pygda_b200.data import Data
    from pygda_b200.models import DGSDA
    g = load_golden("dgsda")["dgsda"]
    est = DGSDA(device="cuda:0", verbose=0, **g["hparams"])
    est.dgsda = est.init_model()
    est.dgsda.load_state_dict(g["state"])
    est.dgsda.train()
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    torch.manual_seed(g["seed"])                             # MMD indices from the CPU generator
    loss, s_logits = est.forward_model(src, tgt)
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    n = 0
    for k, p in est.dgsda.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 2e-4, "grad " + k)
            n += 1
    assert n == len(g["grads"])


def test_dgsda_fit_predict():
    from pygda_b200.models import DGSDA
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1500, 12000, 48, 3, seed=4)
    torch.manual_seed(0)
    model = DGSDA(in_dim=48, hid_dim=32, num_classes=3, K=5, dropout=0.2, epoch=4, lr=0.01, device="cuda:0", verbose=0)
    model.fit(src, tgt)
    logits, labels = model.predict(tgt)
    assert logits.shape == (1500, 3) and labels.shape == (1500,) and torch.isfinite(logits).all()
    s_logits, _ = model.predict(src, source=True)
    assert s_logits.shape == (1500, 3)
