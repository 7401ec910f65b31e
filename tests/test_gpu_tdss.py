"""TDSS on the GPU path (pygda_b200/models/tdss.py, csrc/smooth.cu) against the reference's own
tdss.py (golden fixtures) and the oracle at larger sizes."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.models import TDSS as OracleTDSS

pytestmark = pytest.mark.gpu


def _graph(n, e, seed):
    from pygda_b200.synthetic import powerlaw_edge_index
    return powerlaw_edge_index(n, e, seed=seed, offset=4.0)


def test_khop_smoothing_bit_exact_golden():
    from pygda_b200 import smooth
    g = load_golden("tdss")
    ei = g["target"]["edge_index"].cuda()
    for k, want in g["smooth_khop"].items():
        got = smooth.khop_edge_index(ei, 50, k)
        assert got.dtype == torch.int64 and torch.equal(got.cpu(), want), f"k={k}"


@pytest.mark.parametrize("n,e,k", [(1, 0, 2), (7, 0, 2), (3000, 20000, 2), (800, 3000, 3), (2000, 9000, 1)])
def test_khop_smoothing_bit_exact_oracle(n, e, k):
    from pygda_b200 import smooth
    ei = _graph(n, e, 5) if e else torch.zeros(2, 0, dtype=torch.int64)
    if e:                                   # a few self loops and duplicates
        ei = torch.cat([ei, torch.tensor([[1, 3, 3], [1, 3, 3]]), ei[:, :5]], 1)
    want, _ = OracleTDSS(8, 8, 2, smooth_mode="K-hop", k=k).smoothness(ei, None, n)
    got = smooth.khop_edge_index(ei.cuda(), n, k)
    assert torch.equal(got.cpu(), want)


def test_khop_rejects_bad_index():
    from pygda_b200 import smooth
    from pygda_b200._lib import GdaError
    with pytest.raises(GdaError):
        smooth.khop_edge_index(torch.tensor([[0, 9], [1, 2]]).cuda(), 5, 2)


def test_random_walk_smoothing_properties():
    from pygda_b200 import smooth
    n, length = 400, 4
    ei = _graph(n, 2400, 9)
    ei = ei[:, ei[0] != 7]                  # node 7 has no out-edge: its walk stays put
    out = smooth.random_walk_edge_index(ei.cuda(), n, length, seed=123).cpu()
    again = smooth.random_walk_edge_index(ei.cuda(), n, length, seed=123).cpu()
    other = smooth.random_walk_edge_index(ei.cuda(), n, length, seed=124).cpu()
    assert torch.equal(out, again) and not torch.equal(out, other)
    key = out[0] * n + out[1]
    assert bool((key[1:] > key[:-1]).all()), "row-major, no duplicates (dense_to_sparse order)"
    pairs = set(map(tuple, out.t().tolist()))
    assert all((i, i) in pairs for i in range(n)), "walk[:, 0] = start"
    per_start = torch.bincount(out[1], minlength=n)
    assert int(per_start.max()) <= length + 1 and int(per_start[7]) == 1
    # every visited node is within `length` out-hops of its start
    adj = torch.zeros(n, n, dtype=torch.bool)
    adj[ei[0], ei[1]] = True
    reach = torch.eye(n, dtype=torch.bool)
    frontier = reach.clone()
    for _ in range(length):
        frontier = (frontier.float() @ adj.float()) > 0
        reach |= frontier
    assert bool(reach[out[1], out[0]].all())
    # uniform choice among out-neighbours: first steps from a hub spread over its neighbourhood
    firsts = set()
    hub = int(torch.bincount(ei[0], minlength=n).argmax())
    for s in range(40):
        o = smooth.random_walk_edge_index(ei.cuda(), n, 1, seed=s).cpu()
        firsts |= set(o[0][(o[1] == hub) & (o[0] != hub)].tolist())
    assert len(firsts) >= 10


def test_laplacian_loss_golden_and_larger():
    from pygda_b200 import smooth
    g = load_golden("tdss")["laplacian"]
    f = g["features"].cuda().requires_grad_(True)
    loss = smooth.laplacian_loss(f, g["edge_index"].cuda())
    (2.5 * loss).backward()
    assert_close(loss, g["loss"], 1e-5, "laplacian loss (golden)")
    assert_close(f.grad, 2.5 * g["grad"], 1e-5, "laplacian grad (golden)")
    n, h = 3000, 128
    ei = smooth.khop_edge_index(_graph(n, 20000, 5).cuda(), n, 2)
    x = torch.randn(n, h)
    xc = x.clone().requires_grad_(True)
    ref = OracleTDSS.compute_laplacian_loss(xc, ei.cpu())
    ref.backward()
    xg = x.cuda().requires_grad_(True)
    got = smooth.laplacian_loss(xg, ei)
    got.backward()
    assert_close(got, ref, 1e-4, "laplacian loss (two-hop graph, H=128)")
    assert_close(xg.grad, xc.grad, 1e-4, "laplacian grad")
    # a constant D^-1/2-scaled field has zero loss: f = sqrt(deg) * c on a symmetric graph
    deg = torch.bincount(ei[0], minlength=n).float().cuda()
    flat = deg.sqrt().view(-1, 1).repeat(1, 8).contiguous()
    assert float(smooth.laplacian_loss(flat, ei)) <= 1e-3 * float(flat.square().sum())


def test_forward_model_matches_reference_golden():
    from pygda_b200.data import Data
    from pygda_b200.models import TDSS
    g = load_golden("tdss")
    est = TDSS(device="cuda:0", verbose=0, **g["hparams"])
    est.a2gnn = est.init_model()
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.train()
    src = Data(**g["source"]).to("cuda:0")
    tgt = Data(edge_index_smooth=g["smooth_khop"][2], **g["target"]).to("cuda:0")
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha_grl"])
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    for k, p in est.a2gnn.named_parameters():
        assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)


@pytest.mark.parametrize("mode", ["K-hop", "RW"])
def test_fit_predict(mode, capsys):
    from pygda_b200.models import TDSS
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1500, 9000, 64, 4, seed=2, target_nodes=1200, target_edges=7000)
    torch.manual_seed(0)
    model = TDSS(in_dim=64, hid_dim=32, num_classes=4, smooth_mode=mode, num_layers=2, dropout=0.1, s_pnums=0,
                 t_pnums=3, k=2, rw_len=4, alpha=0.01, beta=1e-4, lr=0.01, epoch=5, device="cuda:0", verbose=0)
    model.fit(src, tgt)
    assert "after smoothness" in capsys.readouterr().out          # tdss.py:505
    assert tgt.edge_index_smooth.is_cuda and tgt.edge_index_smooth.shape[0] == 2
    logits, labels = model.predict(tgt)
    assert logits.shape == (1200, 4) and labels.shape == (1200,) and torch.isfinite(logits).all()
