"""The oracle's restatement of the re-weighted conv family and StruRW (oracle/nn.py: GCNReweight, GSReweight,
ReweightGNN, MixUpGCNConv, MixupBase; oracle/models.py: StruRW) against vectors made by executing the reference's
own pygda/nn/reweight_gnn.py, mixup_gcnconv.py, mixup_base.py and pygda/models/strurw.py (tests/golden/strurw.pt),
and against dense fp64 algebra."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import nn as ONN
from oracle.data import Data
from oracle.models import StruRW


def _layer(name):
    return {"gs_mean": lambda: ONN.GSReweight(10, 7, "mean"), "gs_add": lambda: ONN.GSReweight(10, 7, "add"),
            "gs_mean_normalized": lambda: ONN.GSReweight(10, 7, "mean", normalize_embedding=True),
            "gcn_mean": lambda: ONN.GCNReweight(10, 7, "mean"), "gcn_add": lambda: ONN.GCNReweight(10, 7, "add")}[name]()


@pytest.mark.parametrize("name", ["gs_mean", "gs_add", "gs_mean_normalized", "gcn_mean", "gcn_add"])
def test_reweight_layers_reproduce_the_reference(name):
    g = load_golden("strurw")["layers"]
    c = g["cases"][name]
    layer = _layer(name)
    layer.load_state_dict(c["state"])
    x = g["x"].clone().requires_grad_(True)
    y = layer(x, g["edge_index"], g["edge_weight"], g["lmda"])
    y.backward(c["gout"])
    assert_close(y, c["y"], 1e-6, name)
    assert_close(x.grad, c["gx"], 1e-6, "input gradient")
    for k, p in layer.named_parameters():
        assert_close(p.grad, c["grads"][k], 1e-5, "grad " + k)


def test_gcn_reweight_equals_dense_algebra():
    """'mean': out[r] = (1 / #edges of r) sum_{e: row_e = r} d_row^-1/2 d_col^-1/2 ((1-l) + l rw_e) (x W^T)[col_e] + b with
    d = in-degree at edge_index[1] (gcn_norm's default flow), fp64."""
    g = load_golden("strurw")["layers"]
    c = g["cases"]["gcn_mean"]
    n, ei, rw, l = g["num_nodes"], g["edge_index"], g["edge_weight"].double(), g["lmda"]
    deg = torch.zeros(n, dtype=torch.float64).index_add_(0, ei[1], torch.ones(ei.size(1), dtype=torch.float64))
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    val = dis[ei[0]] * dis[ei[1]] * ((1 - l) + l * rw)
    cnt = torch.zeros(n, dtype=torch.float64).index_add_(0, ei[0], torch.ones(ei.size(1), dtype=torch.float64))
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((ei[0], ei[1]), val, accumulate=True)
    A = A / cnt.clamp(min=1).view(-1, 1)
    ref = A @ (g["x"].double() @ c["state"]["lin.weight"].double().t()) + c["state"]["bias"].double()
    assert_close(c["y"], ref, 1e-5, "reference vector vs dense algebra")
    assert int((cnt == 0).sum()) >= 3                      # the graph has rows without edges


def test_mixup_conv_reproduces_the_reference():
    g = load_golden("strurw")["layers"]
    c = g["cases"]["mixup_conv"]
    conv = ONN.MixUpGCNConv(10, 7)
    conv.load_state_dict(c["state"])
    x = g["x"].clone().requires_grad_(True)
    xc = c["x_cen"].clone().requires_grad_(True)
    y = conv(x, xc, g["edge_index"], g["edge_weight"], g["lmda"])
    y.backward(c["gout"])
    assert_close(y, c["y"], 1e-6, "MixUpGCNConv")
    assert_close(x.grad, c["gx"], 1e-6, "gx")
    assert_close(xc.grad, c["gx_cen"], 1e-6, "gx_cen")
    for k, p in conv.named_parameters():
        assert_close(p.grad, c["grads"][k], 1e-5, "grad " + k)


@pytest.mark.parametrize("name", ["gs", "gcn", "gcn_add", "gs_bn"])
def test_reweight_gnn_reproduces_the_reference(name):
    g = load_golden("strurw")
    c = g["nets"][name]
    L = g["layers"]
    net = ONN.ReweightGNN(**c["hparams"])
    net.load_state_dict(c["state"])
    net.train()
    data = Data(x=L["x"], edge_index=L["edge_index"], edge_weight=L["edge_weight"])
    feat, logits = net(data, data.x)
    assert_close(feat, c["feat"], 1e-6, "features")
    assert_close(logits, c["logits"], 1e-5, "logits")
    (feat * c["gfeat"]).sum().add((logits * c["glogits"]).sum()).backward()
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(c["grads"])
    for k, v in got.items():
        assert_close(v, c["grads"][k], 2e-5, "grad " + k)


@pytest.mark.parametrize("layers", [2, 3])
def test_mixup_base_reproduces_the_reference(layers):
    g = load_golden("strurw")
    c = g["nets"][f"mixup{layers}"]
    L = g["layers"]
    net = ONN.MixupBase(**c["hparams"])
    net.load_state_dict(c["state"])
    net.train()
    feat = net.feat_bottleneck(L["x"], L["edge_index"], c["edge_index_b"], c["lam"], c["perm"].numpy(), L["edge_weight"])
    logits = net.feat_classifier(feat)
    assert_close(feat, c["feat"], 1e-6, "features")
    assert_close(logits, c["logits"], 1e-6, "logits")
    (logits * c["glogits"]).sum().backward()
    for k, p in net.named_parameters():
        assert_close(p.grad, c["grads"][k], 2e-5, "grad " + k)


def test_cal_reweight_is_exact():
    g = load_golden("strurw")["reweight"]
    est = StruRW(in_dim=12, hid_dim=8, num_classes=g["num_classes"])
    s, t = Data(**g["source"]), Data(**g["target"])
    est.cal_reweight(s, t, g["target_pred"])
    assert s.edge_weight.dtype == torch.float32
    assert torch.equal(s.edge_weight, g["edge_weight"])                   # bit-exact
    assert s.edge_weight.unique().numel() > 4


@pytest.mark.parametrize("mode", ["erm", "adv", "mmd", "mixup"])
def test_strurw_forward_model_reproduces_the_reference(mode):
    g = load_golden("strurw")["strurw"]
    r = g["runs"][mode]
    est = StruRW(**r["hparams"])
    est.gnn.load_state_dict(r["state"])
    est.gnn.train()
    mods = [est.gnn]
    if mode == "adv":
        est.domain_discriminator.load_state_dict(r["disc_state"])
        mods.append(est.domain_discriminator)
    s, t = Data(**g["source"]), Data(**g["target"])
    s.edge_weight = torch.ones(s.edge_index.size(1))
    t.edge_weight = torch.ones(t.edge_index.size(1))
    torch.manual_seed(r["seed"])
    np.random.seed(r["np_seed"])
    if mode == "mixup":
        loss, s_logits, t_logits = est.forward_model_mixup(s, t, r["epoch"])
    else:
        loss, s_logits, t_logits = est.forward_model(s, t, r["alpha"], r["epoch"])
    assert torch.equal(s.edge_weight, r["edge_weight"])                   # the re-weighting fired, exactly
    assert_close(loss, r["loss"], 1e-5, "loss")
    assert_close(s_logits, r["source_logits"], 1e-5, "source logits")
    assert_close(t_logits, r["target_logits"], 1e-5, "target logits")
    for m in mods:
        m.zero_grad()
    loss.backward()
    got = {k: p.grad for k, p in est.gnn.named_parameters() if p.grad is not None}
    assert set(got) == set(r["grads"])
    for k, v in got.items():
        assert_close(v, r["grads"][k], 1e-4, "grad " + k)
    if mode == "adv":
        for k, p in est.domain_discriminator.named_parameters():
            assert_close(p.grad, r["disc_grads"][k], 1e-4, "disc grad " + k)
