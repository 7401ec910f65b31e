"""StruRW and the re-weighted conv family on the GPU (pygda_b200/nn/reweight_gnn.py, mixup_gcnconv.py, mixup_base.py,
pygda_b200/models/strurw.py) against vectors made by executing the reference's own pygda/nn/reweight_gnn.py,
mixup_gcnconv.py, mixup_base.py and pygda/models/strurw.py (tests/golden/strurw.pt) and, at a larger size, against
the oracle.  fp32 bar 1e-4 relative (measured values are printed by assert_close on failure); the re-weighted edge
weights are compared bit for bit."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _layer(name):
    from pygda_b200.nn import GCN_reweight, GS_reweight
    return {"gs_mean": lambda: GS_reweight(10, 7, "mean"), "gs_add": lambda: GS_reweight(10, 7, "add"),
            "gs_mean_normalized": lambda: GS_reweight(10, 7, "mean", normalize_embedding=True),
            "gcn_mean": lambda: GCN_reweight(10, 7, "mean"), "gcn_add": lambda: GCN_reweight(10, 7, "add")}[name]()


@pytest.mark.parametrize("name", ["gs_mean", "gs_add", "gs_mean_normalized", "gcn_mean", "gcn_add"])
def test_reweight_layers_golden(name):
    g = load_golden("strurw")["layers"]
    c = g["cases"][name]
    layer = _layer(name).cuda()
    layer.load_state_dict(c["state"])
    x = g["x"].cuda().requires_grad_(True)
    y = layer(x, g["edge_index"].cuda(), g["edge_weight"].cuda(), g["lmda"])
    y.backward(c["gout"].cuda())
    assert_close(y, c["y"], 1e-5, name)
    assert_close(x.grad, c["gx"], 1e-5, "input gradient")
    for k, p in layer.named_parameters():
        assert_close(p.grad, c["grads"][k], 1e-4, "grad " + k)


def test_mixup_conv_golden():
    from pygda_b200.nn import MixUpGCNConv
    g = load_golden("strurw")["layers"]
    c = g["cases"]["mixup_conv"]
    conv = MixUpGCNConv(10, 7).cuda()
    conv.load_state_dict(c["state"])
    x = g["x"].cuda().requires_grad_(True)
    xc = c["x_cen"].cuda().requires_grad_(True)
    y = conv(x, xc, g["edge_index"].cuda(), g["edge_weight"].cuda(), g["lmda"])
    y.backward(c["gout"].cuda())
    assert_close(y, c["y"], 1e-5, "MixUpGCNConv")
    assert_close(x.grad, c["gx"], 1e-5, "gx")
    assert_close(xc.grad, c["gx_cen"], 1e-5, "gx_cen")
    for k, p in conv.named_parameters():
        assert_close(p.grad, c["grads"][k], 1e-4, "grad " + k)


@pytest.mark.parametrize("name", ["gs", "gcn", "gcn_add", "gs_bn"])
def test_reweight_gnn_golden(name):
    from pygda_b200.data import Data
    from pygda_b200.nn import ReweightGNN
    g = load_golden("strurw")
    c, L = g["nets"][name], g["layers"]
    net = ReweightGNN(**c["hparams"]).cuda()
    net.load_state_dict(c["state"])                       # strict: same names, incl. the shared prop_hidden aliases
    net.train()
    data = Data(x=L["x"], edge_index=L["edge_index"], edge_weight=L["edge_weight"]).to("cuda:0")
    feat, logits = net(data, data.x)
    assert_close(feat, c["feat"], 1e-5, "features")
    assert_close(logits, c["logits"], 1e-5, "logits")
    (feat * c["gfeat"].cuda()).sum().add((logits * c["glogits"].cuda()).sum()).backward()
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(c["grads"])
    for k, v in got.items():
        if name == "gs_bn" and k == "mlp_classify.0.bias":
            # a bias in front of a BatchNorm has an analytically ZERO gradient (the batch mean is subtracted):
            # both sides hold only rounding noise there (the reference's CPU run: 1.2e-4; torch's cuDNN BatchNorm
            # backward on the B200: 8.8e-4, GPUTEST_r01), i.e. the noise of a sum of N terms that cancel.  It is
            # bounded by 1e-3 of the scale of the gradient that flows through the same rows (the weight gradient).
            scale = float(c["grads"]["mlp_classify.0.weight"].abs().max())
            assert float(c["grads"][k].abs().max()) <= 1e-3 * scale          # the golden is noise as well
            assert float(v.abs().max()) <= 1e-3 * scale, float(v.abs().max())
            continue
        assert_close(v, c["grads"][k], 1e-4, "grad " + k)


@pytest.mark.parametrize("layers", [2, 3])
def test_mixup_base_golden(layers):
    from pygda_b200.nn import MixupBase
    g = load_golden("strurw")
    c, L = g["nets"][f"mixup{layers}"], g["layers"]
    net = MixupBase(**c["hparams"]).cuda()
    net.load_state_dict(c["state"])
    net.train()
    feat = net.feat_bottleneck(L["x"].cuda(), L["edge_index"].cuda(), c["edge_index_b"].cuda(), c["lam"],
                               c["perm"].numpy(), L["edge_weight"].cuda())
    logits = net.feat_classifier(feat)
    assert_close(feat, c["feat"], 1e-5, "features")
    assert_close(logits, c["logits"], 1e-5, "logits")
    (logits * c["glogits"].cuda()).sum().backward()
    for k, p in net.named_parameters():
        assert_close(p.grad, c["grads"][k], 1e-4, "grad " + k)


def test_cal_reweight_bit_exact_on_the_device():
    from pygda_b200.data import Data
    from pygda_b200.models import StruRW
    g = load_golden("strurw")["reweight"]
    est = StruRW(in_dim=12, hid_dim=8, num_classes=g["num_classes"], device="cuda:0")
    s, t = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    est.cal_reweight(s, t, g["target_pred"].cuda())
    assert s.edge_weight.is_cuda and s.edge_weight.dtype == torch.float32
    assert torch.equal(s.edge_weight.cpu(), g["edge_weight"])


@pytest.mark.parametrize("mode", ["erm", "adv", "mmd", "mixup"])
def test_strurw_forward_model_golden(mode):
    from pygda_b200.data import Data
    from pygda_b200.models import StruRW
    from pygda_b200.nn.layers import Linear
    g = load_golden("strurw")["strurw"]
    r = g["runs"][mode]
    est = StruRW(device="cuda:0", verbose=0, **r["hparams"])
    est.gnn = est.init_model()
    est.gnn.load_state_dict(r["state"])
    est.gnn.train()
    mods = [est.gnn]
    if mode == "adv":
        est.domain_discriminator = Linear(8, 2).cuda()
        est.domain_discriminator.load_state_dict(r["disc_state"])
        mods.append(est.domain_discriminator)
    s, t = est._to_device(Data(**g["source"])), est._to_device(Data(**g["target"]))
    torch.manual_seed(r["seed"])                         # MMD sample indices from the CPU generator
    np.random.seed(r["np_seed"])                         # mixup: beta draw + node shuffle
    if mode == "mixup":
        loss, s_logits, t_logits = est.forward_model_mixup(s, t, r["epoch"])
    else:
        loss, s_logits, t_logits = est.forward_model(s, t, r["alpha"], r["epoch"])
    assert torch.equal(s.edge_weight.cpu(), r["edge_weight"])           # the re-weighting fired; exact
    assert_close(loss, r["loss"], 1e-4, "loss")
    assert_close(s_logits, r["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, r["target_logits"], 1e-4, "target logits")
    loss.backward()
    got = {k: p.grad for k, p in est.gnn.named_parameters() if p.grad is not None}
    assert set(got) == set(r["grads"])
    for k, v in got.items():
        assert_close(v, r["grads"][k], 2e-4, "grad " + k)
    if mode == "adv":
        for k, p in est.domain_discriminator.named_parameters():
            assert_close(p.grad, r["disc_grads"][k], 2e-4, "disc grad " + k)


@pytest.mark.parametrize("backbone,pooling", [("GS", "mean"), ("GCN", "mean"), ("GS", "add")])
def test_reweight_gnn_at_scale_against_the_oracle(backbone, pooling):
    """5 000 nodes / 60 000 edges with hub rows, H = 128 (the 16-byte-per-lane aggregation kernels), directed edge
    list so that some rows are empty."""
    from oracle import nn as ONN
    from oracle.data import Data as OData
    from pygda_b200.data import Data
    from pygda_b200.nn import ReweightGNN
    from pygda_b200.synthetic import powerlaw_edge_index
    n, f, h, c = 5000, 96, 128, 5
    ei = powerlaw_edge_index(n, 60000, seed=8, offset=2.0)[:, :45000]
    torch.manual_seed(2)
    x = torch.randn(n, f)
    ew = torch.rand(ei.size(1)) * 2
    hp = dict(input_dim=f, gnn_dim=h, output_dim=c, cls_dim=64, gnn_layers=2, cls_layers=2, backbone=backbone,
              pooling=pooling, dropout=0.0, rw_lmda=0.8)
    ora = ONN.ReweightGNN(**hp)
    ora.train()
    net = ReweightGNN(**hp).cuda()
    net.load_state_dict(ora.state_dict())
    net.train()
    fr, lr_ = ora(OData(x=x, edge_index=ei, edge_weight=ew), x)
    gl = torch.randn_like(lr_)
    (lr_ * gl).sum().backward()
    d = Data(x=x, edge_index=ei, edge_weight=ew).to("cuda:0")
    fg, lg = net(d, d.x)
    (lg * gl.cuda()).sum().backward()
    assert_close(fg, fr, 1e-4, "features")
    assert_close(lg, lr_, 1e-4, "logits")
    ref = dict(ora.named_parameters())
    for k, p in net.named_parameters():
        if ref[k].grad is not None:
            assert_close(p.grad, ref[k].grad, 2e-4, "grad " + k)


def test_reweighted_graph_is_cached_per_edge_weight_version():
    from pygda_b200.nn.reweight_gnn import message_graph
    ei = torch.tensor([[0, 1, 2, 2], [1, 2, 0, 1]]).cuda()
    w = torch.ones(4).cuda()
    g1 = message_graph(ei, w, 0.8, 3, False, "mean")
    assert message_graph(ei, w, 0.8, 3, False, "mean") is g1
    w.mul_(2.0)                                          # in-place update bumps the version: rebuilt
    g2 = message_graph(ei, w, 0.8, 3, False, "mean")
    assert g2 is not g1
    _, _, v = g2.csr()
    # factor (1 - 0.8) + 0.8 * 2 = 1.8 per edge; 'mean' divides by the edge count of edge_index[0]'s node (node 2
    # has two edges): every one of the three rows sums to 1.8
    assert_close(v.sum(), torch.tensor(3 * 1.8), 1e-6, "sum of folded weights")


@pytest.mark.parametrize("mode,gnn", [("erm", "GS"), ("adv", "GCN"), ("mmd", "GS"), ("mixup", "GS")])
def test_strurw_fit_predict(mode, gnn, capsys):
    from pygda_b200.models import StruRW
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1200, 9000, 40, 3, seed=6)
    torch.manual_seed(0)
    np.random.seed(0)
    model = StruRW(in_dim=40, hid_dim=32, num_classes=3, num_layers=2, cls_dim=16, dropout=0.2, gnn=gnn, mode=mode,
                   ew_start=2, ew_freq=2, epoch=5, lr=0.01, device="cuda:0", verbose=2)
    model.fit(src, tgt)
    out = capsys.readouterr().out
    assert out.count("edge reweight...") == 2            # epochs 1 and 3
    assert out.count("Epoch") >= 5 or out.count("epoch") >= 5
    logits, labels = model.predict(tgt)
    assert logits.shape == (1200, 3) and labels.shape == (1200,) and torch.isfinite(logits).all()


def test_lincomb_kernels():
    """ops.add / ops.lerp2 (gda_scale_f32 + gda_axpy_f32) forward and backward against torch."""
    from pygda_b200 import ops
    torch.manual_seed(1)
    a0, b0 = torch.randn(1000, 37), torch.randn(1000, 37)
    for fn, ref, wa, wb in ((lambda a, b: ops.lerp2(a, b, 0.3), lambda a, b: a * 0.3 + b * (1 - 0.3), 0.3, 0.7),
                            (ops.add, lambda a, b: a + b, 1.0, 1.0)):
        a, b = a0.clone().cuda().requires_grad_(True), b0.clone().cuda().requires_grad_(True)
        y = fn(a, b)
        go = torch.randn_like(y)
        y.backward(go)
        assert_close(y, ref(a0, b0), 1e-6, "value")
        assert_close(a.grad, wa * go, 1e-6, "grad a")
        assert_close(b.grad, wb * go, 1e-6, "grad b")
