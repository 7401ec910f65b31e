"""Golden vectors for the GRAPH-LEVEL mode of A2GNN, UDAGCN and GRADE (global_mean_pool call sites
pygda/nn/a2gnn_base.py:140-141, pygda/models/udagcn.py:169-170, pygda/nn/grade_base.py:154-157), made by EXECUTING THE
REFERENCE'S OWN FILES on two small collated graph batches:

    python tests/golden/make_golden_graph_mode.py        # build container only (needs /root/reference)

Writes tests/golden/graph_mode.pt: per estimator the state_dict, loss, source / target logits and every parameter
gradient of one ``forward_model``; plus the A2GNN quirk that ``adv=True`` cannot run in graph mode (node-count labels
against pooled features, a2gnn.py:200-204): the exception type and message of the reference."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference, REPO  # noqa: E402

sys.path.insert(0, REPO)
from oracle.data import Data, collate_graphs  # noqa: E402
from pygda_b200.synthetic import graph_dataset  # noqa: E402


def pack(d):
    return {k: getattr(d, k) for k in ("x", "edge_index", "y", "batch", "num_graphs") if getattr(d, k, None) is not None}


def randomise_vectors(mod):
    with torch.no_grad():
        for p in mod.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)


def main():
    ref = load_reference()
    gsrc = collate_graphs([Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in graph_dataset(14, 9, 2.0, 6, 3, seed=5)])
    gtgt = collate_graphs([Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in graph_dataset(11, 11, 3.0, 6, 3, seed=6)])
    out = {"source": pack(gsrc), "target": pack(gtgt)}

    # ---- A2GNN (a2gnn.py:146-213), graph mode: pooled bottleneck, Linear classifier, MMD over graphs ----
    hp = dict(in_dim=6, hid_dim=16, num_classes=3, mode="graph", num_layers=2, dropout=0.0, s_pnums=0, t_pnums=3,
              adv=False, weight=2.0)
    torch.manual_seed(61)
    est = ref.a2gnn.A2GNN(device="cpu", **hp)
    est.a2gnn = est.init_model()
    randomise_vectors(est.a2gnn)
    est.a2gnn.train()
    state = {k: v.clone() for k, v in est.a2gnn.state_dict().items()}
    torch.manual_seed(63)                                    # MMD sample indices
    loss, s_logits, t_logits = est.forward_model(gsrc, gtgt, 0.2)
    est.a2gnn.zero_grad()
    loss.backward()
    out["a2gnn"] = {"hparams": hp, "alpha": 0.2, "seed": 63, "state": state, "loss": loss.detach().clone(),
                    "source_logits": s_logits.detach().clone(), "target_logits": t_logits.detach().clone(),
                    "grads": {k: p.grad.clone() for k, p in est.a2gnn.named_parameters() if p.grad is not None}}
    hp_adv = dict(hp, adv=True)
    torch.manual_seed(61)
    est = ref.a2gnn.A2GNN(device="cpu", **hp_adv)
    est.a2gnn = est.init_model()
    try:
        est.forward_model(gsrc, gtgt, 0.2)
        out["a2gnn_adv_error"] = None
    except Exception as exc:                                 # noqa: BLE001
        out["a2gnn_adv_error"] = {"type": type(exc).__name__, "message": str(exc), "hparams": hp_adv}

    # ---- UDAGCN (udagcn.py:131-201), graph mode: encodings pooled before the heads ----
    hp = dict(in_dim=6, hid_dim=16, num_classes=3, mode="graph", num_layers=2, ppmi=False, adv_dim=10, epoch=300)
    torch.manual_seed(65)
    est = ref.udagcn.UDAGCN(device="cpu", **hp)
    est.udagcn = est.init_model()
    for m in est.udagcn.models:
        randomise_vectors(m)
        m.eval()
    # the encoder's dropout layers are a plain list and stay active (udagcn_base.py:47): switch them off by hand
    for d in est.udagcn.encoder.dropout_layers:
        d.p = 0.0
    state = {k: v.clone() for k, v in est.udagcn.state_dict().items()}
    loss, s_logits, t_logits = est.forward_model(gsrc, gtgt, 0.04, 120)
    est.udagcn.zero_grad()
    loss.backward()
    out["udagcn"] = {"hparams": hp, "alpha": 0.04, "epoch": 120, "state": state, "loss": loss.detach().clone(),
                     "source_logits": s_logits.detach().clone(), "target_logits": t_logits.detach().clone(),
                     "grads": {k: p.grad.clone() for k, p in est.udagcn.named_parameters() if p.grad is not None}}

    # ---- GRADE (grade.py:129-197), graph mode: per-layer pooled features, labels per GRAPH ----
    for disc in ("JS", "MMD"):
        hp = dict(in_dim=6, hid_dim=16, num_classes=3, mode="graph", num_layers=2, dropout=0.0, disc=disc, weight=0.5)
        torch.manual_seed(67)
        est = ref.grade.GRADE(device="cpu", **hp)
        est.grade = est.init_model()
        randomise_vectors(est.grade)
        est.grade.train()
        state = {k: v.clone() for k, v in est.grade.state_dict().items()}
        torch.manual_seed(69)
        loss, s_logits, t_logits = est.forward_model(gsrc, gtgt, 0.3)
        est.grade.zero_grad()
        loss.backward()
        out["grade_" + disc.lower()] = {
            "hparams": hp, "alpha": 0.3, "seed": 69, "state": state, "loss": loss.detach().clone(),
            "source_logits": s_logits.detach().clone(), "target_logits": t_logits.detach().clone(),
            "grads": {k: p.grad.clone() for k, p in est.grade.named_parameters() if p.grad is not None}}

    torch.save(out, os.path.join(HERE, "graph_mode.pt"))
    print("wrote graph_mode.pt", os.path.getsize(os.path.join(HERE, "graph_mode.pt")), "bytes; adv error:",
          out["a2gnn_adv_error"] and (out["a2gnn_adv_error"]["type"], out["a2gnn_adv_error"]["message"][:80]))


if __name__ == "__main__":
    main()
