"""End-to-end parity of the A2GNN estimator with the reference's own forward_model
(golden fixtures) and with the oracle at a larger size; size-independent properties at
the benchmark scale."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.data import Data as OData
from oracle.models import A2GNN as OracleA2GNN

pytestmark = pytest.mark.gpu


def _estimator(h):
    from pygda_b200.models import A2GNN
    est = A2GNN(device="cuda:0", verbose=0, **h)
    est.a2gnn = est.init_model()
    return est


@pytest.mark.parametrize("name", ["a2gnn_mmd", "a2gnn_adv"])
def test_forward_model_matches_reference_golden(name):
    from pygda_b200.data import Data
    g = load_golden(name)
    est = _estimator(g["hparams"])
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.train()
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    torch.manual_seed(g["seed"])            # forward_model draws the MMD indices itself
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"])
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    assert torch.equal(s_logits.argmax(1).cpu(), g["source_logits"].argmax(1))
    for k, p in est.a2gnn.named_parameters():
        assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)


def test_train_steps_track_the_oracle():
    """3 optimiser steps (dropout=0) from the same weights: losses and weights agree."""
    from pygda_b200.data import Data
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import domain_pair
    from oracle import mmd as OM
    h = dict(in_dim=200, hid_dim=64, num_classes=5, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=5,
             adv=False, weight=10, weight_decay=0.005, lr=0.01, epoch=200)
    src, tgt = domain_pair(3000, 30000, 200, 5, seed=3, target_nodes=2500, target_edges=24000)
    torch.manual_seed(0)
    ora = OracleA2GNN(device="cpu", **h)
    ora.mmd_sqdist = lambda z: OM.pairwise_sqdist_blocked(z, 250)
    est = _estimator(h)
    est.a2gnn.load_state_dict(ora.a2gnn.state_dict())
    opt = Adam(est.a2gnn.parameters(), lr=h["lr"], weight_decay=h["weight_decay"])
    osrc, otgt = OData(x=src.x, edge_index=src.edge_index, y=src.y), OData(x=tgt.x, edge_index=tgt.edge_index, y=tgt.y)
    for step in range(3):
        torch.manual_seed(100 + step)
        idx = OM.draw_mmd_indices(3000, 2500)
        ora.mmd_indices = idx
        ref_loss, ref_s, ref_t = ora.train_step(osrc, otgt, epoch=step)
        loss, s_logits, t_logits, _ = est.train_step(src, tgt, est.alpha_at(step, h["epoch"]), opt,
                                                     mmd_indices=idx)
        assert_close(loss, torch.tensor(ref_loss), 2e-4, f"loss step {step}")
        assert_close(s_logits, ref_s, 2e-4, f"source logits step {step}")
        assert_close(t_logits, ref_t, 2e-4, f"target logits step {step}")
    for (k, p), (_, q) in zip(est.a2gnn.named_parameters(), ora.a2gnn.named_parameters()):
        assert_close(p, q, 1e-3, "weights after 3 steps: " + k)


def test_fit_predict_api_and_training_learns():
    from pygda_b200.models import A2GNN
    from pygda_b200.data import Data

    def homophilous(n, seed):
        """labels = strongest of 4 feature groups; edges only between same-label nodes, so the
        one propagation step in the classifier (a2gnn_base.py:171-174) keeps the class signal."""
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(n, 64, generator=g)
        y = x.view(n, 4, -1).sum(-1).argmax(1)
        order = torch.argsort(y)
        counts = torch.bincount(y, minlength=4).tolist()
        eis, off = [], 0
        for c in counts:
            u = order[off + torch.randint(c, (4 * c,), generator=g)]
            v = order[off + torch.randint(c, (4 * c,), generator=g)]
            eis.append(torch.stack([torch.cat([u, v]), torch.cat([v, u])]))
            off += c
        return Data(x=x, edge_index=torch.cat(eis, 1), y=y)

    src, tgt = homophilous(2000, 1), homophilous(2000, 2)
    torch.manual_seed(0)
    model = A2GNN(in_dim=64, hid_dim=32, num_classes=4, num_layers=2, dropout=0.1, s_pnums=0, t_pnums=3,
                  weight=1, lr=0.01, epoch=80, device="cuda:0", verbose=0)
    model.fit(src, tgt)                           # host tensors: copied to the device every step
    logits, labels = model.predict(tgt)
    assert logits.shape == (2000, 4) and labels.shape == (2000,) and logits.is_cuda
    logits_s, labels_s = model.predict(None, source=True)   # `data` is ignored, like the reference
    acc = (logits_s.argmax(1) == labels_s).float().mean().item()
    assert acc > 0.6, acc


def test_first_layer_sharing_is_value_preserving():
    from pygda_b200.data import Data
    g = load_golden("a2gnn_mmd")
    est = _estimator(g["hparams"])
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.eval()
    tgt = Data(**g["target"]).to("cuda:0")
    a = est.a2gnn(tgt, 3)
    b = est.a2gnn(tgt, 3, first_layer=est.a2gnn.first_conv(tgt.x, tgt.edge_index, 3))
    assert torch.equal(a, b)


def test_benchmark_scale_properties():
    """Config-2 shaped target graph (100k nodes / 1M edges, H=128): properties that do not
    need the CPU oracle -- linearity, A_hat 1-eigenvector, transpose adjointness."""
    from pygda_b200 import ops
    from pygda_b200.graph import Graph
    from pygda_b200.synthetic import powerlaw_edge_index
    n, h = 100_000, 128
    ei = powerlaw_edge_index(n, 1_000_000, seed=1).cuda()
    gr = Graph(ei, n)
    assert gr.nnz == 1_100_000
    x, y = torch.randn(n, h, device="cuda"), torch.randn(n, h, device="cuda")
    ax, ay = ops.spmm(gr, x), ops.spmm(gr, y)
    assert_close(ops.spmm(gr, 2 * x - 3 * y), 2 * ax - 3 * ay, 1e-5, "linearity")
    # <A x, y> == <x, A^T y>
    lhs = (ax.double() * y.double()).sum()
    rhs = (x.double() * ops.spmm(gr, y, transpose=True).double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * abs(float(lhs)) + 1e-3
    # D^-1/2 (A+I) D^-1/2 has eigenvector sqrt(deg) with eigenvalue 1 (symmetric graph)
    deg = torch.bincount(ei[1], minlength=n).float() + 1
    v = deg.sqrt().view(-1, 1).repeat(1, 4).contiguous()
    assert_close(ops.spmm(gr, v), v, 1e-5, "eigenvector")
