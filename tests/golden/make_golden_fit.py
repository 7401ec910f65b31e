"""Training TRAJECTORIES made by running the reference's own ``fit`` / ``predict`` (pygda/models/a2gnn.py:215-411,
pygda/models/strurw.py:325-444, 669-700) for a few epochs on small graphs -- loaders, alpha schedule, optimiser
configuration and update rule, epoch loop and predict, end to end:

    python tests/golden/make_golden_fit.py        # build container only (needs /root/reference)

Writes tests/golden/fit.pt: per run the initial state_dict (captured from the ``init_model`` call inside ``fit``), the
torch CPU RNG state right after it (the MMD indices of the following epochs are drawn from that stream), the final
state_dict, and ``predict`` on the target (and source) graph.  dropout = 0: the dropout streams cannot be matched."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference, REPO  # noqa: E402

sys.path.insert(0, REPO)
from make_golden import small_graph  # noqa: E402
from oracle.data import Data  # noqa: E402


def capture_init(est, attr_box, post=None):
    """Wrap est.init_model so that the state it creates inside fit() and the RNG state after it are recorded."""
    real = est.init_model

    def wrapped(**kw):
        net = real(**kw)
        if post is not None:
            post(net)
        attr_box["state"] = {k: v.clone() for k, v in net.state_dict().items()}
        attr_box["rng_state"] = torch.get_rng_state().clone()
        return net
    est.init_model = wrapped


def main():
    ref = load_reference()
    src, tgt = small_graph(60, 220, 20, 3, seed=21), small_graph(52, 180, 20, 3, seed=22, with_loops=False)
    blob = {"source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
            "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y}, "runs": {}}

    for name, kw in (("a2gnn_mmd", dict(adv=False, weight=3)), ("a2gnn_adv", dict(adv=True, weight=2))):
        hp = dict(in_dim=20, hid_dim=12, num_classes=3, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=4, lr=0.01,
                  weight_decay=0.005, epoch=4, **kw)
        torch.manual_seed(71)
        est = ref.a2gnn.A2GNN(device="cpu", verbose=0, **hp)
        box = {}
        capture_init(est, box)
        est.fit(Data(**blob["source"]), Data(**blob["target"]))
        t_logits, t_labels = est.predict(Data(**blob["target"]))
        s_logits, s_labels = est.predict(Data(**blob["source"]), source=True)
        # predict() ignores its ``data`` argument and re-iterates the loaders stored by fit()
        other_logits, other_labels = est.predict(Data(**blob["source"]))          # source graph, source=False
        blob["runs"][name] = {"hparams": hp, "init_state": box["state"], "rng_state": box["rng_state"],
                              "final_state": {k: v.clone() for k, v in est.a2gnn.state_dict().items()},
                              "target_logits": t_logits.clone(), "target_labels": t_labels.clone(),
                              "source_logits": s_logits.clone(), "source_labels": s_labels.clone(),
                              "predict_ignores_data": bool(torch.equal(other_logits, t_logits) and
                                                           torch.equal(other_labels, t_labels))}

    # StruRW, GS backbone, 'erm' objective; the edge re-weighting fires in epochs 1 and 3 (and, PyG's loaders building a
    # new batch object per epoch, lasts for that epoch's source pass only)
    hp = dict(in_dim=20, hid_dim=12, num_classes=3, num_layers=2, cls_dim=8, cls_layers=2, dropout=0.0, gnn="GS",
              pooling="mean", ew_start=2, ew_freq=2, lamb=0.8, mode="erm", lr=0.01, weight_decay=0.001, epoch=4)
    torch.manual_seed(73)
    est = ref.strurw.StruRW(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box)
    est.fit(Data(edge_weight=None, **blob["source"]), Data(edge_weight=None, **blob["target"]))   # PyG: missing -> None
    t = Data(**blob["target"])
    t.edge_weight = torch.ones(t.edge_index.size(1))
    t_logits, t_labels = est.predict(t)
    blob["runs"]["strurw_erm"] = {"hparams": hp, "init_state": box["state"], "rng_state": box["rng_state"],
                                  "final_state": {k: v.clone() for k, v in est.gnn.state_dict().items()},
                                  "target_logits": t_logits.clone(), "target_labels": t_labels.clone()}
    def finish(name, est, net_attr, hp, box, predict_kw=()):
        net = getattr(est, net_attr)
        run = {"hparams": hp, "init_state": box["state"], "rng_state": box["rng_state"],
               "final_state": {k: v.clone() for k, v in net.state_dict().items()}}
        t_logits, t_labels = est.predict(Data(**blob["target"]))
        run["target_logits"], run["target_labels"] = t_logits.clone(), t_labels.clone()
        if "source" in predict_kw:
            s_logits, s_labels = est.predict(Data(**blob["source"]), source=True)
            run["source_logits"], run["source_labels"] = s_logits.clone(), s_labels.clone()
        blob["runs"][name] = run

    # UDAGCN (udagcn.py:203-308), adjacency view only; the encoder's never-registered dropout layers (udagcn_base.py:47)
    # are switched off by hand, in the network fit() has just built
    def no_encoder_dropout(net):
        for d in net.encoder.dropout_layers:
            d.p = 0.0
    hp = dict(in_dim=20, hid_dim=12, num_classes=3, num_layers=2, ppmi=False, adv_dim=8, lr=0.01, weight_decay=0.003,
              epoch=4)
    torch.manual_seed(75)
    est = ref.udagcn.UDAGCN(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box, post=no_encoder_dropout)
    # domain_model holds an nn.Dropout(0.1) that IS registered and active in train mode: its draws come from the CPU
    # generator on the reference side and cannot be matched -> p = 0 as well
    real = est.init_model

    def with_plain_domain_model(**kw):
        net = real(**kw)
        for m in net.domain_model:
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        return net
    est.init_model = with_plain_domain_model
    est.fit(Data(**blob["source"]), Data(**blob["target"]))
    finish("udagcn", est, "udagcn", hp, box, ("source",))

    for disc in ("JS", "MMD"):
        hp = dict(in_dim=20, hid_dim=12, num_classes=3, num_layers=2, dropout=0.0, disc=disc, weight=0.5, lr=0.01,
                  weight_decay=0.01, epoch=4)
        torch.manual_seed(77)
        est = ref.grade.GRADE(device="cpu", verbose=0, **hp)
        box = {}
        capture_init(est, box)
        est.fit(Data(**blob["source"]), Data(**blob["target"]))
        finish("grade_" + disc.lower(), est, "grade", hp, box, ("source",))

    for backbone in ("gcn", "gat"):
        hp = dict(in_dim=20, hid_dim=12, num_classes=3, num_layers=2, dropout=0.0, gnn=backbone, lr=0.02,
                  weight_decay=0.001, epoch=4)
        torch.manual_seed(79)
        est = ref.gnn.GNN(device="cpu", verbose=0, **hp)
        box = {}
        capture_init(est, box)
        est.fit(Data(**blob["source"]), Data(**blob["target"]))
        finish("gnn_" + backbone, est, "gnn", hp, box)

    hp = dict(in_dim=20, hid_dim=12, num_classes=3, mode="node", smooth_mode="K-hop", num_layers=2, dropout=0.0,
              s_pnums=0, t_pnums=3, k=2, alpha=0.5, beta=0.05, lr=0.01, weight_decay=0.005, epoch=4)
    torch.manual_seed(81)
    est = ref.tdss.TDSS(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box)
    tt = Data(edge_attr=None, **blob["target"])
    est.fit(Data(**blob["source"]), tt)
    finish("tdss", est, "a2gnn", hp, box, ("source",))

    hp = dict(in_dim=20, hid_dim=12, num_classes=3, K=4, alpha=0.05, beta=0.5, gamma=0.05, dropout=0.0, lr=0.01,
              weight_decay=0.0005, epoch=4)
    torch.manual_seed(83)
    est = ref.dgsda.DGSDA(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box)
    est.fit(Data(**blob["source"]), Data(**blob["target"]))
    finish("dgsda", est, "dgsda", hp, box, ("source",))

    # AdaGCN (adagcn.py:200-319), node level: 10 critic iterations with WGAN-GP per step.  The critic is created inside
    # fit() right after init_model (:264-276) from the same CPU generator, so the recorded RNG state reproduces its
    # initial weights too; its nn.Dropout(0.1) -- and the encoder's own nn.Dropout(0.1), adagcn_base.py:59 -- draw from that
    # generator in train mode and are built with p = 0 here.
    hp = dict(in_dim=20, hid_dim=12, num_classes=3, mode="node", num_layers=2, dropout=0.0, adv_dim=8, gp_weight=5,
              domain_weight=0.5, lr=0.01, weight_decay=0.001, epoch=3)
    real_dropout = torch.nn.Dropout

    class NoDropout(real_dropout):
        def __init__(self, p=0.5, inplace=False):
            super().__init__(0.0, inplace)
    torch.manual_seed(85)
    est = ref.adagcn.AdaGCN(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box)
    torch.nn.Dropout = NoDropout
    try:
        est.fit(Data(**blob["source"]), Data(**blob["target"]))
    finally:
        torch.nn.Dropout = real_dropout
    finish("adagcn_node", est, "adagcn", hp, box, ("source",))
    blob["runs"]["adagcn_node"]["critic_final_state"] = {k: v.clone() for k, v in est.discriminator.state_dict().items()}

    # StruRW with the adversarial objective (GS backbone) and with MMD on the GCN backbone
    for name, extra in (("strurw_adv", dict(gnn="GS", mode="adv")), ("strurw_mmd", dict(gnn="GCN", mode="mmd")),
                        ("strurw_mixup", dict(mode="mixup"))):
        hp = dict(in_dim=20, hid_dim=12, num_classes=3, num_layers=2, cls_dim=8, cls_layers=2, dropout=0.0,
                  pooling="mean", ew_start=2, ew_freq=2, lamb=0.8, lr=0.01, weight_decay=0.001, epoch=4, **extra)
        torch.manual_seed(91)
        est = ref.strurw.StruRW(device="cpu", verbose=0, **hp)
        box = {}
        capture_init(est, box)
        import numpy as np
        np.random.seed(95)                                   # mixup: beta draw + node shuffle per step (strurw.py:301,723)
        est.fit(Data(edge_weight=None, **blob["source"]), Data(edge_weight=None, **blob["target"]))
        t = Data(**blob["target"])
        t.edge_weight = torch.ones(t.edge_index.size(1))
        t_logits, t_labels = est.predict(t)
        run = {"np_seed": 95, "hparams": hp, "init_state": box["state"], "rng_state": box["rng_state"],
               "final_state": {k: v.clone() for k, v in est.gnn.state_dict().items()},
               "target_logits": t_logits.clone(), "target_labels": t_labels.clone()}
        if extra["mode"] == "adv":
            run["disc_final_state"] = {k: v.clone() for k, v in est.domain_discriminator.state_dict().items()}
        blob["runs"][name] = run

    # What fit() PRINTS with verbose=2 (utils/utility.py): per epoch the summed loss and the micro-F1 of the source
    # predictions -- A2GNN scores the TRAINING-mode logits of the step (a2gnn.py:321-329), GNN and StruRW re-predict in eval
    # mode after every step (gnn.py:219-226, strurw.py:423-431).  Same seeds as the runs above, so the same trajectories.
    import contextlib
    import io
    import re

    def logged(make_est, seed, fit_args):
        torch.manual_seed(seed)
        est = make_est()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            est.fit(*fit_args())
        rows = re.findall(r"Epoch (\d+): loss ([-0-9.]+), source acc ([0-9.]+)", buf.getvalue())
        return [(int(e), float(l), float(a)) for e, l, a in rows]

    plain = lambda: (Data(**blob["source"]), Data(**blob["target"]))                                   # noqa: E731
    weighted = lambda: (Data(edge_weight=None, **blob["source"]), Data(edge_weight=None, **blob["target"]))   # noqa: E731
    blob["runs"]["a2gnn_mmd"]["log"] = logged(
        lambda: ref.a2gnn.A2GNN(device="cpu", verbose=2, **blob["runs"]["a2gnn_mmd"]["hparams"]), 71, plain)
    blob["runs"]["gnn_gcn"]["log"] = logged(
        lambda: ref.gnn.GNN(device="cpu", verbose=2, **blob["runs"]["gnn_gcn"]["hparams"]), 79, plain)
    blob["runs"]["strurw_erm"]["log"] = logged(
        lambda: ref.strurw.StruRW(device="cpu", verbose=2, **blob["runs"]["strurw_erm"]["hparams"]), 73, weighted)

    # Graph-level mode with shuffled mini-batches (a2gnn.py:266-286, grade.py:214-252): DataLoader(batch_size=8,
    # shuffle=True) over lists of small graphs -- the batch order comes from torch's sampler on the CPU generator.
    from pygda_b200.synthetic import graph_dataset
    gs = [Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in graph_dataset(21, 9, 2.0, 6, 3, seed=7)]
    gt = [Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in graph_dataset(19, 11, 3.0, 6, 3, seed=8)]
    blob["graph_source"] = [{"x": d.x, "edge_index": d.edge_index, "y": d.y} for d in gs]
    blob["graph_target"] = [{"x": d.x, "edge_index": d.edge_index, "y": d.y} for d in gt]

    def finish_graph(name, est, net_attr, hp, box):
        net = getattr(est, net_attr)
        blob["runs"][name] = {"hparams": hp, "init_state": box["state"], "rng_state": box["rng_state"],
                              "final_state": {k: v.clone() for k, v in net.state_dict().items()}}

    hp = dict(in_dim=6, hid_dim=12, num_classes=3, mode="graph", num_layers=2, dropout=0.0, s_pnums=0, t_pnums=2,
              adv=False, weight=2.0, lr=0.01, weight_decay=0.005, epoch=3, batch_size=8)
    torch.manual_seed(87)
    est = ref.a2gnn.A2GNN(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box)
    est.fit(gs, gt)
    finish_graph("a2gnn_graph", est, "a2gnn", hp, box)

    hp = dict(in_dim=6, hid_dim=12, num_classes=3, mode="graph", num_layers=2, dropout=0.0, disc="JS", weight=0.5,
              lr=0.01, weight_decay=0.01, epoch=3, batch_size=8)
    torch.manual_seed(89)
    est = ref.grade.GRADE(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box)
    est.fit(gs, gt)
    finish_graph("grade_graph", est, "grade", hp, box)

    # UDAGCN in graph mode, batch_size=0 (one shuffled batch of ALL graphs per epoch).  Pins the CachedGCNConv quirk end to
    # end: the normalised graph is cached per cache_name at the FIRST batch (cached_gcn_conv.py:132-136) and silently
    # re-used for the differently shuffled batches of the later epochs.
    hp = dict(in_dim=6, hid_dim=12, num_classes=3, mode="graph", num_layers=2, ppmi=False, adv_dim=8, lr=0.01,
              weight_decay=0.003, epoch=3, batch_size=0)
    torch.manual_seed(93)
    est = ref.udagcn.UDAGCN(device="cpu", verbose=0, **hp)
    box = {}
    capture_init(est, box, post=no_encoder_dropout)
    real_u = est.init_model

    def plain_domain_model_u(**kw):
        net = real_u(**kw)
        for m in net.domain_model:
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        return net
    est.init_model = plain_domain_model_u
    est.fit(gs, gt)
    finish_graph("udagcn_graph", est, "udagcn", hp, box)

    torch.save(blob, os.path.join(HERE, "fit.pt"))
    print("wrote fit.pt", os.path.getsize(os.path.join(HERE, "fit.pt")), "bytes")


if __name__ == "__main__":
    main()
