"""The CUDA-graph replay of the training-loop body (pygda_b200/models/graphed.py) against the
eager step it was captured from: same kernels, same numbers."""
import copy

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu

H = dict(in_dim=200, hid_dim=128, num_classes=5, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=4,
         weight=10, weight_decay=0.005, lr=0.01, epoch=200)


def _pair():
    from pygda_b200.synthetic import domain_pair
    return domain_pair(3000, 30000, 200, 5, seed=3, target_nodes=2500, target_edges=24000, device="cuda:0")


def _est(adv, dropout=0.0):
    from pygda_b200.models import A2GNN
    from pygda_b200.optim import Adam
    h = dict(H, adv=adv, dropout=dropout)
    est = A2GNN(device="cuda:0", verbose=0, **h)
    torch.manual_seed(0)
    est.a2gnn = est.init_model()
    opt = Adam(est.a2gnn.parameters(), lr=h["lr"], weight_decay=h["weight_decay"])
    return est, opt


@pytest.mark.parametrize("adv", [False, True])
def test_replay_equals_eager_steps(adv):
    from pygda_b200.models.graphed import GraphedStep
    src, tgt = _pair()
    steps = 4
    # eager arm
    est, opt = _est(adv)
    state0 = copy.deepcopy(est.a2gnn.state_dict())
    torch.manual_seed(11)                                  # MMD indices come from the CPU generator
    eager = []
    for i in range(steps):
        loss, s_logits, t_logits, _ = est.train_step(src, tgt, est.alpha_at(i, 200), opt)
        eager.append((loss.item(), s_logits.clone(), t_logits.clone()))
    eager_params = [p.detach().clone() for p in est.a2gnn.parameters()]
    # graphed arm: step 0 is the eager warm-up inside the constructor, steps 1.. are replays
    est2, opt2 = _est(adv)
    est2.a2gnn.load_state_dict(state0)
    torch.manual_seed(11)
    g = GraphedStep(est2, src, tgt, opt2, warmup=1, alpha_fn=lambda i: est2.alpha_at(i, 200))
    assert g.launches_per_replay > 20
    got = [tuple(t.clone() for t in g.warmup_results[0])]
    for i in range(1, steps):
        loss, s_logits, t_logits = g(est2.alpha_at(i, 200))
        got.append((loss.clone(), s_logits.clone(), t_logits.clone()))
    torch.cuda.synchronize()
    for i, ((l0, s0, t0), (l1, s1, t1)) in enumerate(zip(eager, got)):
        # not bit-identical: the column-sum / CE / MMD reductions accumulate with float atomics (arrival order)
        assert abs(l1.item() - l0) <= 1e-5 * abs(l0) + 1e-7, f"loss step {i}: {l1.item()} vs {l0}"
        assert_close(s1, s0, 1e-5, f"source logits step {i}")
        assert_close(t1, t0, 1e-5, f"target logits step {i}")
    for p, q in zip(est2.a2gnn.parameters(), eager_params):
        # Adam divides by sqrt(v): where a gradient element is ~0 the last-bit differences of the atomically
        # accumulated reductions are amplified to a visible fraction of lr (same bar as __graft_entry__.smoke)
        assert_close(p, q, 1e-3, "weights after replays")


def test_replays_draw_fresh_dropout_masks_and_indices():
    from pygda_b200.models.graphed import GraphedStep
    from pygda_b200.utils.mmd import draw_indices
    src, tgt = _pair()
    est, opt = _est(False, dropout=0.5)
    for p in est.a2gnn.parameters():
        p.requires_grad_(True)
    opt.lr = 0.0                                           # frozen weights: only masks / indices change
    opt.weight_decay = 0.0
    g = GraphedStep(est, src, tgt, opt, warmup=1)
    idx = draw_indices(3000, 2500)
    a = g(mmd_indices=idx)[2].clone()
    b = g(mmd_indices=idx)[2].clone()
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert not torch.equal(a, b), "dropout masks must differ between replays"
    # staged indices are what the kernels read
    assert torch.equal(g.s_idx.cpu(), idx[0]) and torch.equal(g.t_idx.cpu(), idx[1])


def test_fit_with_cuda_graph_learns():
    from pygda_b200.models import A2GNN
    src, tgt = _pair()
    torch.manual_seed(0)
    model = A2GNN(in_dim=200, hid_dim=32, num_classes=5, num_layers=2, dropout=0.1, s_pnums=0, t_pnums=3,
                  weight=1, lr=0.01, epoch=30, device="cuda:0", verbose=0)
    model.cuda_graph = True
    model.fit(src.to("cpu"), tgt.to("cpu"))
    assert model.graphed_step.replays == 29
    logits, labels = model.predict(tgt)
    assert logits.shape == (2500, 5) and torch.isfinite(logits).all()
