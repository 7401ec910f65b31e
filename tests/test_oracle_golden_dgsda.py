"""The oracle's DGSDA restatement (oracle/nn.py: BernProp, DGSDABase; oracle/models.py: DGSDA; upstream
get_laplacian / add_self_loops in oracle/pyg_ops.py) against vectors made by executing the reference's own
pygda/nn/dgsda_base.py and pygda/models/dgsda.py (tests/golden/dgsda.pt), and against dense fp64 algebra."""
from math import comb

import torch

from conftest import assert_close, load_golden
from oracle import nn as ONN
from oracle import pyg_ops as P
from oracle.data import Data
from oracle.models import DGSDA


def test_bernprop_reproduces_the_reference():
    g = load_golden("dgsda")["bernprop"]
    for K, c in g["cases"].items():
        prop = ONN.BernProp(K)
        with torch.no_grad():
            prop.temp.copy_(c["temp"])
        x = c["x"].clone().requires_grad_(True)
        y = prop(x, g["edge_index"])
        y.backward(c["gout"])
        assert_close(y, c["y"], 1e-6, f"BernProp K={K}")
        assert_close(x.grad, c["gx"], 1e-6, "input gradient")
        assert_close(prop.temp.grad, c["gtemp"], 1e-5, "temp gradient")


def test_bernprop_equals_the_dense_bernstein_polynomial():
    """sum_k C(K,k)/2^K relu(t_k) L^k (2I - L)^(K-k) x with L = I - D^-1/2 A D^-1/2 in fp64 (self loops removed,
    multi-edges counted, propagation = transpose of the (row, col) matrix)."""
    g = load_golden("dgsda")["bernprop"]
    n, ei = g["num_nodes"], g["edge_index"]
    keep = ei[0] != ei[1]
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((ei[0][keep], ei[1][keep]), torch.ones(int(keep.sum()), dtype=torch.float64), accumulate=True)
    dis = A.sum(1).pow(-0.5)
    dis[torch.isinf(dis)] = 0
    L = torch.eye(n, dtype=torch.float64) - dis[:, None] * A * dis[None, :]
    Lt, Mt = L.t(), (2 * torch.eye(n, dtype=torch.float64) - L).t()      # propagate: out[col] += w * x[row]
    c = g["cases"][4]
    x, t = c["x"].double(), torch.relu(c["temp"].double())
    ref = sum(comb(4, k) / 2 ** 4 * t[k] * torch.linalg.matrix_power(Lt, k) @ torch.linalg.matrix_power(Mt, 4 - k) @ x
              for k in range(5))
    assert_close(c["y"], ref, 1e-5, "reference vector vs dense algebra")
    ei1, w1 = P.get_laplacian(ei, None, "sym", torch.float64, n)
    dense = torch.zeros(n, n, dtype=torch.float64)
    dense.index_put_((ei1[0], ei1[1]), w1, accumulate=True)
    assert_close(dense, L, 1e-12, "get_laplacian")


def test_dgsda_forward_model_reproduces_the_reference():
    g = load_golden("dgsda")["dgsda"]
    est = DGSDA(**g["hparams"])
    est.dgsda.load_state_dict(g["state"])
    est.dgsda.train()
    torch.manual_seed(g["seed"])
    loss, s_logits = est.forward_model(Data(**g["source"]), Data(**g["target"]))
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    est.dgsda.zero_grad()
    loss.backward()
    n = 0
    for k, p in est.dgsda.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)
            n += 1
    assert n == len(g["grads"]) and n >= 7
