"""Oracle TDSS (oracle/models.py) against the fixtures made by executing the reference's own
pygda/models/tdss.py (tests/golden/make_golden.py), plus dense fp64 pins of the upstream ops it
restates (spspmm pattern, coalesce) -- SURVEY.md section 8(f) row 1."""
import torch

from conftest import assert_close, load_golden
from oracle import pyg_ops as P
from oracle.data import Data
from oracle.models import TDSS


def test_two_hop_and_khop_smoothing_bit_exact():
    g = load_golden("tdss")
    ei = g["target"]["edge_index"]
    assert torch.equal(TDSS.two_hop(ei, 50), g["two_hop"])
    for k, want in g["smooth_khop"].items():
        est = TDSS(24, 16, 4, smooth_mode="K-hop", k=k)
        got, _ = est.smoothness(ei, None, 50)
        assert torch.equal(got, want), f"k={k}"


def test_two_hop_pattern_is_dense_a_squared():
    g = load_golden("tdss")
    ei = g["target"]["edge_index"]
    a = torch.zeros(50, 50, dtype=torch.float64)
    a[ei[0], ei[1]] = 1.0
    two = ((a @ a) > 0) & ~torch.eye(50, dtype=torch.bool)
    want = (two | (a > 0)).nonzero().t()                  # row-major = coalesced order
    assert torch.equal(TDSS.two_hop(ei, 50), want)


def test_random_walk_smoothing_construction():
    g = load_golden("tdss")["smooth_rw"]
    ei = load_golden("tdss")["target"]["edge_index"]
    est = TDSS(24, 16, 4, smooth_mode="RW", rw_len=4)
    torch.manual_seed(g["seed"])
    got, _ = est.smoothness(ei, None, 50)
    assert torch.equal(got, g["edge_index"])
    # every edge (v, i) has v on the walk from i; walks follow existing edges or stay put
    walk = g["walk"]
    assert walk.shape == (50, 5) and torch.equal(walk[:, 0], torch.arange(50))
    have = set(map(tuple, ei.t().tolist()))
    for i in range(50):
        for t in range(4):
            a, b = int(walk[i, t]), int(walk[i, t + 1])
            assert (a, b) in have or a == b
    pairs = set(map(tuple, got.t().tolist()))
    assert pairs == {(int(walk[i, t]), i) for i in range(50) for t in range(5)}


def test_laplacian_loss_golden_and_dense():
    g = load_golden("tdss")["laplacian"]
    f = g["features"].clone().requires_grad_(True)
    loss = TDSS.compute_laplacian_loss(f, g["edge_index"])
    loss.backward()
    assert_close(loss, g["loss"], 1e-6, "laplacian loss")
    assert_close(f.grad, g["grad"], 1e-6, "laplacian grad")
    # dense fp64: 1/2 sum_e |g_r - g_c|^2 with g = D^-1/2 f, D = out-degree (edge multiplicity counted)
    ei = g["edge_index"]
    fd = g["features"].double()
    deg = torch.zeros(50, dtype=torch.float64).index_add_(0, ei[0], torch.ones(ei.size(1), dtype=torch.float64))
    gg = fd * deg.pow(-0.5).masked_fill(deg == 0, 0).view(-1, 1)
    dense = 0.5 * ((gg[ei[0]] - gg[ei[1]]) ** 2).sum()
    assert_close(loss, dense, 1e-5, "laplacian loss vs fp64")


def test_forward_model_golden():
    g = load_golden("tdss")
    est = TDSS(device="cpu", **g["hparams"])
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.train()
    est.mmd_indices = (g["source_idx"], g["target_idx"])
    src = Data(**g["source"])
    tgt = Data(edge_index_smooth=g["smooth_khop"][2], **g["target"])
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha_grl"])
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(s_logits, g["source_logits"], 1e-5, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-5, "target logits")
    for k, p in est.a2gnn.named_parameters():
        assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)
