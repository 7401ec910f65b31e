"""Training trajectories: the oracle's loop body (oracle/models.py ``train_step``) + loaders (oracle/data.py) run for
the reference's number of epochs from the reference's initial weights and RNG state must land on the weights and the
``predict`` outputs the reference's own ``fit`` / ``predict`` produced (tests/golden/fit.pt,
tests/golden/make_golden_fit.py)."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.data import Data, FullBatchNeighborLoader
from oracle.models import A2GNN, StruRW


@pytest.mark.parametrize("name", ["a2gnn_mmd", "a2gnn_adv"])
def test_a2gnn_fit_trajectory(name):
    G = load_golden("fit")
    r = G["runs"][name]
    hp = r["hparams"]
    est = A2GNN(**hp)
    est.a2gnn.load_state_dict(r["init_state"])
    torch.set_rng_state(r["rng_state"])
    src_loader, tgt_loader = FullBatchNeighborLoader(Data(**G["source"])), FullBatchNeighborLoader(Data(**G["target"]))
    for epoch in range(hp["epoch"]):
        for s, t in zip(src_loader, tgt_loader):
            est.train_step(s, t, epoch)
    for k, v in est.a2gnn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-5, "weights after fit: " + k)
    t_logits, t_labels = est.predict(next(iter(tgt_loader)))
    s_logits, s_labels = est.predict(next(iter(src_loader)), source=True)
    assert_close(t_logits, r["target_logits"], 1e-5, "predict(target)")
    assert_close(s_logits, r["source_logits"], 1e-5, "predict(source)")
    assert torch.equal(t_labels, r["target_labels"]) and torch.equal(s_labels, r["source_labels"])


def test_strurw_fit_trajectory(capsys):
    G = load_golden("fit")
    r = G["runs"]["strurw_erm"]
    hp = r["hparams"]
    est = StruRW(**hp)
    est.gnn.load_state_dict(r["init_state"])
    torch.set_rng_state(r["rng_state"])
    s0, t0 = Data(**G["source"]), Data(**G["target"])
    s0.edge_weight = torch.ones(s0.edge_index.size(1))
    t0.edge_weight = torch.ones(t0.edge_index.size(1))
    src_loader, tgt_loader = FullBatchNeighborLoader(s0), FullBatchNeighborLoader(t0)
    fired = 0
    for epoch in range(hp["epoch"]):
        for s, t in zip(src_loader, tgt_loader):
            est.train_step(s, t, epoch)
            fired += int(not torch.equal(s.edge_weight, torch.ones_like(s.edge_weight)))
    assert fired == 2                                       # epochs 1 and 3; a new batch object per epoch
    for k, v in est.gnn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-5, "weights after fit: " + k)
    t = Data(**G["target"])
    t.edge_weight = torch.ones(t.edge_index.size(1))
    logits, labels = est.predict(t)
    assert_close(logits, r["target_logits"], 1e-5, "predict(target)")
    assert torch.equal(labels, r["target_labels"])
