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


@pytest.mark.parametrize("name", ["strurw_erm", "strurw_adv", "strurw_mmd", "strurw_mixup"])
def test_strurw_fit_trajectory(name, capsys):
    G = load_golden("fit")
    r = G["runs"][name]
    hp = r["hparams"]
    est = StruRW(**hp)
    est.gnn.load_state_dict(r["init_state"])
    torch.set_rng_state(r["rng_state"])
    if hp["mode"] == "adv":                                 # created inside fit() after init_model (strurw.py:364-372)
        est.domain_discriminator = torch.nn.Linear(hp["hid_dim"], 2)
        est.optimizer = torch.optim.Adam(list(est.gnn.parameters()) + list(est.domain_discriminator.parameters()),
                                         lr=hp["lr"], weight_decay=hp["weight_decay"])
    s0, t0 = Data(**G["source"]), Data(**G["target"])
    s0.edge_weight = torch.ones(s0.edge_index.size(1))
    t0.edge_weight = torch.ones(t0.edge_index.size(1))
    src_loader, tgt_loader = FullBatchNeighborLoader(s0), FullBatchNeighborLoader(t0)
    import numpy as np
    np.random.seed(r.get("np_seed", 0))
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
    if hp["mode"] == "adv":
        for k, v in est.domain_discriminator.state_dict().items():
            assert_close(v, r["disc_final_state"][k], 1e-5, "discriminator after fit: " + k)


def _run_loop(est, net, G, r, step, prepare=None):
    net.load_state_dict(r["init_state"])
    torch.set_rng_state(r["rng_state"])
    s0, t0 = Data(**G["source"]), Data(**G["target"])
    if prepare is not None:
        prepare(s0, t0)
    src_loader, tgt_loader = FullBatchNeighborLoader(s0), FullBatchNeighborLoader(t0)
    for epoch in range(r["hparams"]["epoch"]):
        for s, t in zip(src_loader, tgt_loader):
            step(s, t, epoch)
    for k, v in net.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-5, "weights after fit: " + k)
    return next(iter(src_loader)), next(iter(tgt_loader))


def test_udagcn_fit_trajectory():
    from oracle.models import UDAGCN
    G = load_golden("fit")
    r = G["runs"]["udagcn"]
    est = UDAGCN(**r["hparams"])
    est.udagcn.encoder.dropout_layers = [torch.nn.Identity() for _ in est.udagcn.encoder.dropout_layers]
    for m in est.udagcn.domain_model:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    s, t = _run_loop(est, est.udagcn, G, r, lambda a, b, e: est.train_step(a, b, e))
    for m in est.udagcn.models:
        m.eval()
    with torch.no_grad():
        assert_close(est.udagcn.cls_model(est.udagcn.encode(t, "target")), r["target_logits"], 1e-5, "predict(target)")
        assert_close(est.udagcn.cls_model(est.udagcn.encode(s, "source")), r["source_logits"], 1e-5, "predict(source)")


@pytest.mark.parametrize("disc", ["js", "mmd"])
def test_grade_fit_trajectory(disc):
    from oracle.models import GRADE
    G = load_golden("fit")
    r = G["runs"]["grade_" + disc]
    est = GRADE(**r["hparams"])
    s, t = _run_loop(est, est.grade, G, r, lambda a, b, e: est.train_step(a, b, e))
    est.grade.eval()
    with torch.no_grad():
        assert_close(est.grade(t)[0], r["target_logits"], 1e-5, "predict(target)")
        assert_close(est.grade(s)[0], r["source_logits"], 1e-5, "predict(source)")


@pytest.mark.parametrize("backbone", ["gcn", "gat"])
def test_gnn_fit_trajectory(backbone):
    from oracle.models import GNN
    G = load_golden("fit")
    r = G["runs"]["gnn_" + backbone]
    hp = r["hparams"]
    est = GNN(**hp)
    opt = torch.optim.Adam(est.gnn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])   # gnn.py:195-199

    def step(a, b, e):
        est.gnn.train()
        loss, _, _ = est.forward_model(a, b)
        opt.zero_grad()
        loss.backward()
        opt.step()
    s, t = _run_loop(est, est.gnn, G, r, step)
    est.gnn.eval()
    with torch.no_grad():
        assert_close(est.gnn(t.x, t.edge_index), r["target_logits"], 1e-5, "predict(target)")


def test_tdss_fit_trajectory():
    from oracle.models import TDSS
    G = load_golden("fit")
    r = G["runs"]["tdss"]
    est = TDSS(**r["hparams"])

    def prepare(s0, t0):                                    # tdss.py:497-506: done by fit() before the loaders are built
        t0.edge_index_smooth, t0.edge_attr_smooth = est.smoothness(t0.edge_index, None, t0.x.shape[0])
    s, t = _run_loop(est, est.a2gnn, G, r, lambda a, b, e: est.train_step(a, b, e), prepare)
    assert_close(est.predict(t)[0], r["target_logits"], 1e-5, "predict(target)")
    assert_close(est.predict(s, source=True)[0], r["source_logits"], 1e-5, "predict(source)")


def test_dgsda_fit_trajectory():
    from oracle.models import DGSDA
    G = load_golden("fit")
    r = G["runs"]["dgsda"]
    est = DGSDA(**r["hparams"])
    s, t = _run_loop(est, est.dgsda, G, r, lambda a, b, e: est.train_step(a, b))
    est.dgsda.eval()
    with torch.no_grad():
        assert_close(est.dgsda(t, False), r["target_logits"], 1e-5, "predict(target)")
        assert_close(est.dgsda(s, True), r["source_logits"], 1e-5, "predict(source)")


def _critic_from_the_stream(hid, adv):
    """The reference builds its critic inside fit() right after init_model (adagcn.py:264-270): the same constructor
    calls on the same generator state give the same initial weights."""
    return torch.nn.Sequential(torch.nn.Linear(hid, adv), torch.nn.ReLU(), torch.nn.Dropout(0.0),
                               torch.nn.Linear(adv, 1), torch.nn.Sigmoid())


def test_adagcn_fit_trajectory():
    from oracle.models import AdaGCN
    G = load_golden("fit")
    r = G["runs"]["adagcn_node"]
    hp = r["hparams"]
    est = AdaGCN(**hp)
    est.adagcn.load_state_dict(r["init_state"])
    est.adagcn.encoder.dropout.p = 0.0                     # the encoder keeps its own Dropout(0.1) (adagcn_base.py:59,84):
    torch.set_rng_state(r["rng_state"])                    # built with p = 0 on the reference side too
    est.discriminator = _critic_from_the_stream(hp["hid_dim"], hp["adv_dim"])
    est.c_optimizer = torch.optim.Adam(est.discriminator.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    src_loader, tgt_loader = FullBatchNeighborLoader(Data(**G["source"])), FullBatchNeighborLoader(Data(**G["target"]))
    for epoch in range(hp["epoch"]):
        for s, t in zip(src_loader, tgt_loader):
            est.adagcn.train()
            loss, _, _ = est.forward_model(s, t)
            est.optimizer.zero_grad()                      # clears what the critic loop left on the encoder (:302)
            loss.backward()
            est.optimizer.step()
    for k, v in est.adagcn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-5, "encoder after fit: " + k)
    for k, v in est.discriminator.state_dict().items():
        assert_close(v, r["critic_final_state"][k], 1e-5, "critic after fit: " + k)
    est.adagcn.eval()
    with torch.no_grad():
        t = next(iter(tgt_loader))
        assert_close(est.adagcn.cls_model(est.adagcn(t)), r["target_logits"], 1e-5, "predict(target)")


@pytest.mark.parametrize("name", ["a2gnn_graph", "grade_graph"])
def test_graph_mode_minibatch_fit_trajectory(name):
    """Shuffled mini-batches of graphs: the batch order comes from torch's own sampler machinery on the CPU generator
    (PyG's DataLoader is torch's), the zip of the two loaders stops with the shorter one."""
    from oracle.data import GraphDataLoader
    from oracle.models import GRADE
    G = load_golden("fit")
    r = G["runs"][name]
    hp = dict(r["hparams"])
    bs = hp.pop("batch_size")
    est, net = (A2GNN(**hp), None) if name.startswith("a2gnn") else (GRADE(**hp), None)
    net = est.a2gnn if name.startswith("a2gnn") else est.grade
    net.load_state_dict(r["init_state"])
    torch.set_rng_state(r["rng_state"])
    gs = [Data(**d) for d in G["graph_source"]]
    gt = [Data(**d) for d in G["graph_target"]]
    src_loader, tgt_loader = GraphDataLoader(gs, bs, shuffle=True), GraphDataLoader(gt, bs, shuffle=True)
    for epoch in range(hp["epoch"]):
        for s, t in zip(src_loader, tgt_loader):
            est.train_step(s, t, epoch)
    for k, v in net.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-5, "weights after fit: " + k)


def test_udagcn_graph_mode_fit_trajectory_with_the_stale_graph_cache():
    """One shuffled batch of all graphs per epoch: the conv layers keep the normalised graph of the FIRST batch
    (cached_gcn_conv.py:132-136) and apply it to the differently ordered later batches -- as the reference does."""
    from oracle.data import GraphDataLoader
    from oracle.models import UDAGCN
    G = load_golden("fit")
    r = G["runs"]["udagcn_graph"]
    hp = dict(r["hparams"])
    hp.pop("batch_size")
    est = UDAGCN(**hp)
    est.udagcn.load_state_dict(r["init_state"])
    est.udagcn.encoder.dropout_layers = [torch.nn.Identity() for _ in est.udagcn.encoder.dropout_layers]
    for m in est.udagcn.domain_model:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    torch.set_rng_state(r["rng_state"])
    gs = [Data(**d) for d in G["graph_source"]]
    gt = [Data(**d) for d in G["graph_target"]]
    src_loader, tgt_loader = GraphDataLoader(gs, len(gs), shuffle=True), GraphDataLoader(gt, len(gt), shuffle=True)
    orders = []
    for epoch in range(hp["epoch"]):
        for s, t in zip(src_loader, tgt_loader):
            orders.append(s.y.tolist())
            est.train_step(s, t, epoch)
    assert orders[0] != orders[1]                          # the batches really are ordered differently
    for k, v in est.udagcn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-5, "weights after fit: " + k)
