"""GATConv(heads=1, concat=False): oracle pinned against dense fp64 attention; CUDA kernels vs the oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close, load_golden
from oracle import nn as ONN
from oracle.data import Data
from oracle.models import GNN as OracleGNN


def test_oracle_gat_equals_dense_attention():
    torch.manual_seed(0)
    n, f, c = 30, 7, 5
    ei = torch.randint(n, (2, 120))
    conv = ONN.GATConv(f, c).double()
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    x = torch.randn(n, f, dtype=torch.float64)
    h = x @ conv.lin_src.weight.t()
    a_s, a_d = (h * conv.att_src.view(1, c)).sum(1), (h * conv.att_dst.view(1, c)).sum(1)
    keep = ei[0] != ei[1]
    src = torch.cat([ei[0][keep], torch.arange(n)])
    dst = torch.cat([ei[1][keep], torch.arange(n)])
    out = torch.zeros(n, c, dtype=torch.float64)
    for i in range(n):                                     # explicit per-target softmax (multi-edges count twice)
        js = src[dst == i]
        e = F.leaky_relu(a_s[js] + a_d[i], 0.2)
        w = torch.softmax(e, 0)
        out[i] = (w.unsqueeze(1) * h[js]).sum(0)
    assert_close(conv(x, ei), out + conv.bias, 1e-10, "dense attention")


def test_oracle_gnn_gat_golden():
    g = load_golden("gnn_gat")
    est = OracleGNN(**g["hparams"])
    est.gnn.load_state_dict(g["state"])
    est.gnn.train()
    loss, s_logits, _ = est.forward_model(Data(**g["source"]), Data(**g["target"]))
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "logits")


@pytest.mark.gpu
def test_gat_conv_forward_backward_vs_oracle():
    from pygda_b200.nn import GATConv
    from pygda_b200.synthetic import powerlaw_edge_index
    torch.manual_seed(1)
    n, f, c = 4000, 48, 32
    ei = powerlaw_edge_index(n, 40000, seed=5, offset=2.0)
    ei = torch.cat([ei, torch.tensor([[3, 3, 9], [3, 3, 9]])], 1)       # existing self loops, a duplicate
    ref = ONN.GATConv(f, c)
    with torch.no_grad():
        ref.bias.uniform_(-0.5, 0.5)
    conv = GATConv(f, c).cuda()
    conv.load_state_dict(ref.state_dict())
    x = torch.randn(n, f)
    coef = torch.randn(n, c)
    xr = x.clone().requires_grad_(True)
    (ref(xr, ei) * coef).sum().backward()
    xg = x.cuda().requires_grad_(True)
    y = conv(xg, ei.cuda())
    (y * coef.cuda()).sum().backward()
    assert_close(y, ref(x, ei), 1e-4, "fwd")
    assert_close(xg.grad, xr.grad, 1e-4, "dx")
    for (k, p), (_, q) in zip(conv.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, 2e-4, "grad " + k)


@pytest.mark.gpu
def test_gnn_gat_forward_model_golden():
    from pygda_b200.data import Data as GData
    from pygda_b200.models import GNN
    g = load_golden("gnn_gat")
    est = GNN(device="cuda:0", verbose=0, **g["hparams"])
    est.gnn = est.init_model()
    est.gnn.load_state_dict(g["state"])
    est.gnn.train()
    src, tgt = GData(**g["source"]).to("cuda:0"), GData(**g["target"]).to("cuda:0")
    loss, s_logits, t_logits = est.forward_model(src, tgt)
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    for k, p in est.gnn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 2e-4, "grad " + k)
