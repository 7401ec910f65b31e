"""Generate golden input/output vectors by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference package cannot be imported whole (torch_geometric / torch_scatter /
torch_sparse are not installed), so ref_loader.py loads the hot-path files
individually from /root/reference with the minimal stand-ins of
tests/golden/_pyg_stub for the three missing packages.  What is pinned by these
fixtures is therefore the reference's OWN code (gcn_norm, PropGCNConv, A2GNNBase,
A2GNN.forward_model, MMD, GradReverse, CachedGCNConv, ...); the upstream PyG ops
underneath are the restatement in oracle/pyg_ops.py (see oracle/__init__.py).

Outputs: tests/golden/*.pt (small tensors, committed).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference, REPO  # noqa: E402

sys.path.insert(0, REPO)
from pygda_b200.synthetic import citation_graph  # noqa: E402
from oracle.data import Data  # noqa: E402


def small_graph(n, e, f, c, seed, with_loops=True):
    d = citation_graph(n, e, f, c, seed=seed, degree_offset=4.0)
    ei = d.edge_index
    if with_loops:   # a few existing self loops and a duplicate edge: exercises add_remaining_self_loops
        extra = torch.tensor([[1, 3, 3, int(ei[0, 0])], [1, 3, 3, int(ei[1, 0])]])
        ei = torch.cat([ei[:, :7], extra, ei[:, 7:]], 1)
    return Data(x=d.x, edge_index=ei, y=d.y)


def main():
    ref = load_reference()
    torch.manual_seed(1234)
    out = {}

    # ---- gcn_norm (pygda/nn/prop_gcn_conv.py:24-81) ---------------------------------------
    g = small_graph(40, 120, 12, 3, seed=7)
    w = torch.rand(g.edge_index.size(1)) + 0.5
    cases = {}
    for name, kw in {"plain": {}, "improved": {"improved": True}, "weighted": {"edge_weight": w},
                     "noloops": {"add_self_loops": False}}.items():
        ei, ew = ref.prop_gcn_conv.gcn_norm(g.edge_index, kw.get("edge_weight"), 40,
                                            kw.get("improved", False), kw.get("add_self_loops", True),
                                            torch.float32)
        cases[name] = {"edge_index_out": ei, "weight_out": ew}
    out["gcn_norm"] = {"edge_index": g.edge_index, "num_nodes": 40, "edge_weight": w, "cases": cases}

    # ---- CachedGCNConv.norm (pygda/nn/cached_gcn_conv.py:63-103) --------------------------
    ei, ew = ref.cached_gcn_conv.CachedGCNConv.norm(g.edge_index, 40, None, False, torch.float32)
    out["cached_norm"] = {"edge_index": g.edge_index, "num_nodes": 40, "edge_index_out": ei, "weight_out": ew}

    # ---- PropGCNConv.forward for k = 0, 1, 3 (prop_gcn_conv.py:153-215) -------------------
    conv = ref.prop_gcn_conv.PropGCNConv(12, 8)
    with torch.no_grad():
        conv.bias.uniform_(-0.1, 0.1)
    pc = {"state": {k: v.clone() for k, v in conv.state_dict().items()}, "x": g.x,
          "edge_index": g.edge_index, "out": {}, "grad_w": {}, "grad_x": {}}
    for k in (0, 1, 3):
        x = g.x.clone().requires_grad_(True)
        conv.zero_grad()
        y = conv(x, g.edge_index, k)
        (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
        pc["out"][k] = y.detach().clone()
        pc["grad_w"][k] = conv.lin.weight.grad.clone()
        pc["grad_x"][k] = x.grad.clone()
    out["prop_gcn_conv"] = pc

    # ---- CachedGCNConv.forward ------------------------------------------------------------
    cconv = ref.cached_gcn_conv.CachedGCNConv(12, 8)
    with torch.no_grad():
        cconv.bias.uniform_(-0.1, 0.1)
    out["cached_gcn_conv"] = {"state": {k: v.clone() for k, v in cconv.state_dict().items()}, "x": g.x,
                              "edge_index": g.edge_index,
                              "out": cconv(g.x, g.edge_index, "c").detach().clone()}

    # ---- MMD (pygda/utils/mmd.py) -- the real file, no stubs involved --------------------
    s_feat = torch.randn(70, 16)
    t_feat = torch.randn(55, 16) * 1.3 + 0.2
    s_req, t_req = s_feat.clone().requires_grad_(True), t_feat.clone().requires_grad_(True)
    torch.manual_seed(99)
    loss = ref.mmd.MMD(s_req, t_req, sampling_num=64, times=3)
    loss.backward()
    torch.manual_seed(99)   # the same draws, recorded for the kernels (mmd.py:148-149)
    s_idx = torch.randint(70, (3, 64))
    t_idx = torch.randint(55, (3, 64))
    out["mmd"] = {"source": s_feat, "target": t_feat, "seed": 99, "sampling_num": 64, "times": 3,
                  "source_idx": s_idx, "target_idx": t_idx, "loss": loss.detach().clone(),
                  "grad_source": s_req.grad.clone(), "grad_target": t_req.grad.clone(),
                  "get_mmd_full": ref.mmd.get_MMD(s_feat[:50], t_feat[:50]).clone()}

    # ---- GradReverse (pygda/nn/reverse_layer.py) -- real file ------------------------------
    xr = torch.randn(5, 4, requires_grad=True)
    yr = ref.reverse_layer.GradReverse.apply(xr, 0.37)
    yr.backward(torch.ones_like(yr) * 2.0)
    out["grad_reverse"] = {"x": xr.detach().clone(), "alpha": 0.37, "y": yr.detach().clone(),
                           "grad": xr.grad.clone()}

    # ---- A2GNN.forward_model, MMD and adversarial variants (models/a2gnn.py:146-213) ------
    src = small_graph(60, 200, 24, 4, seed=11)
    tgt = small_graph(50, 150, 24, 4, seed=12)
    for adv in (False, True):
        torch.manual_seed(5 + int(adv))
        est = ref.a2gnn.A2GNN(in_dim=24, hid_dim=16, num_classes=4, mode='node', num_layers=2,
                              dropout=0.0, s_pnums=0, t_pnums=3, adv=adv, weight=10, device='cpu')
        est.a2gnn = est.init_model()
        with torch.no_grad():
            for p in est.a2gnn.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.1, 0.1)
        est.a2gnn.train()
        state = {k: v.clone() for k, v in est.a2gnn.state_dict().items()}
        torch.manual_seed(77)
        loss, s_logits, t_logits = est.forward_model(src, tgt, 0.3)
        est.a2gnn.zero_grad()
        loss.backward()
        grads = {k: p.grad.clone() for k, p in est.a2gnn.named_parameters()}
        torch.manual_seed(77)
        s_idx = torch.randint(60, (5, 1000))
        t_idx = torch.randint(50, (5, 1000))
        out["a2gnn_adv" if adv else "a2gnn_mmd"] = {
            "source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
            "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
            "hparams": dict(in_dim=24, hid_dim=16, num_classes=4, num_layers=2, dropout=0.0, s_pnums=0,
                            t_pnums=3, adv=adv, weight=10),
            "alpha": 0.3, "seed": 77, "state": state, "source_idx": s_idx, "target_idx": t_idx,
            "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
            "target_logits": t_logits.detach().clone(), "grads": grads}

    # ---- UDAGCN.forward_model, ppmi=False (models/udagcn.py:131-201) -------------------------
    # The encoder's dropout list is always active in the reference (udagcn_base.py:47); the RNG
    # streams cannot match, so both sides run with those layers replaced by identities and the
    # registered Dropout of the domain model in eval mode.
    torch.manual_seed(21)
    uest = ref.udagcn.UDAGCN(in_dim=24, hid_dim=16, num_classes=4, mode='node', num_layers=2, ppmi=False,
                             adv_dim=10, epoch=300, device='cpu')
    uest.udagcn = uest.init_model()
    uest.udagcn.encoder.dropout_layers = [torch.nn.Identity() for _ in uest.udagcn.encoder.dropout_layers]
    with torch.no_grad():
        for p in uest.udagcn.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    for m in uest.udagcn.models:
        m.eval()
    ustate = {k: v.clone() for k, v in uest.udagcn.state_dict().items()}
    loss, s_logits, t_logits = uest.forward_model(src, tgt, 0.04, 120)
    uest.udagcn.zero_grad()
    loss.backward()
    out["udagcn"] = {
        "source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
        "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
        "hparams": dict(in_dim=24, hid_dim=16, num_classes=4, num_layers=2, ppmi=False, adv_dim=10, epoch=300),
        "alpha": 0.04, "epoch": 120, "state": ustate, "loss": loss.detach().clone(),
        "source_logits": s_logits.detach().clone(), "target_logits": t_logits.detach().clone(),
        "grads": {k: p.grad.clone() for k, p in uest.udagcn.named_parameters() if p.grad is not None}}

    # ---- GRADE.forward_model, disc = JS / MMD / C (models/grade.py:129-197) --------------------
    for disc in ("JS", "MMD", "C"):
        torch.manual_seed(31)
        gest = ref.grade.GRADE(in_dim=24, hid_dim=16, num_classes=4, mode='node', num_layers=2, dropout=0.0,
                               disc=disc, weight=0.5, device='cpu')
        gest.grade = gest.init_model()
        with torch.no_grad():
            for p in gest.grade.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.1, 0.1)
        gest.grade.train()
        gstate = {k: v.clone() for k, v in gest.grade.state_dict().items()}
        torch.manual_seed(55)
        loss, s_logits, t_logits = gest.forward_model(src, tgt, 0.3)
        gest.grade.zero_grad()
        loss.backward()
        torch.manual_seed(55)
        s_idx = torch.randint(50, (5, 1000))      # mind = min(60, 50) rows of each side (grade.py:177-182)
        t_idx = torch.randint(50, (5, 1000))
        out["grade_" + disc.lower()] = {
            "source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
            "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
            "hparams": dict(in_dim=24, hid_dim=16, num_classes=4, num_layers=2, dropout=0.0, disc=disc, weight=0.5),
            "alpha": 0.3, "seed": 55, "state": gstate, "source_idx": s_idx, "target_idx": t_idx,
            "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
            "target_logits": t_logits.detach().clone(),
            "grads": {k: p.grad.clone() for k, p in gest.grade.named_parameters() if p.grad is not None}}

    # ---- AdaGCN.forward_model (models/adagcn.py:138-198): 10 critic iterations with WGAN-GP, node and graph mode
    from pygda_b200.synthetic import graph_dataset
    from oracle.data import collate_graphs
    gsrc = collate_graphs([Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in graph_dataset(12, 9, 2.0, 6, 2, seed=1)])
    gtgt = collate_graphs([Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in graph_dataset(9, 11, 3.0, 6, 2, seed=2)])
    for mode, (a, b), dims in (("node", (src, tgt), (24, 4)), ("graph", (gsrc, gtgt), (6, 2))):
        torch.manual_seed(41)
        aest = ref.adagcn.AdaGCN(in_dim=dims[0], hid_dim=16, num_classes=dims[1], mode=mode, num_layers=2, adv_dim=10,
                                 gp_weight=5, domain_weight=0.5, lr=0.01, weight_decay=0.001, device='cpu')
        aest.adagcn = aest.init_model()
        aest.discriminator = torch.nn.Sequential(                      # as created inside fit (:264-270)
            torch.nn.Linear(16, 10), torch.nn.ReLU(), torch.nn.Dropout(0.1), torch.nn.Linear(10, 1),
            torch.nn.Sigmoid())
        aest.c_optimizer = torch.optim.Adam(aest.discriminator.parameters(), lr=0.01, weight_decay=0.001)
        with torch.no_grad():
            for p in aest.adagcn.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.1, 0.1)
        aest.adagcn.eval()             # dropout off on both sides (RNG streams cannot match)
        aest.discriminator.eval()
        astate = {k: v.clone() for k, v in aest.adagcn.state_dict().items()}
        dstate = {k: v.clone() for k, v in aest.discriminator.state_dict().items()}
        torch.manual_seed(43)          # gradient_penalty draws torch.rand on the CPU generator (:405,:415,:421)
        loss, s_logits, t_logits = aest.forward_model(a, b)
        aest.adagcn.zero_grad()
        loss.backward()
        def pack(d):
            return {k: getattr(d, k) for k in ("x", "edge_index", "y", "batch", "num_graphs") if getattr(d, k, None) is not None}
        out["adagcn_" + mode] = {
            "source": pack(a), "target": pack(b), "seed": 43,
            "hparams": dict(in_dim=dims[0], hid_dim=16, num_classes=dims[1], mode=mode, num_layers=2, adv_dim=10,
                            gp_weight=5, domain_weight=0.5, lr=0.01, weight_decay=0.001),
            "state": astate, "critic_state": dstate,
            "critic_state_after": {k: v.clone() for k, v in aest.discriminator.state_dict().items()},
            "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
            "target_logits": t_logits.detach().clone(),
            "grads": {k: p.grad.clone() for k, p in aest.adagcn.named_parameters() if p.grad is not None}}

    # ---- GNN (gcn / gat backbone) forward_model (models/gnn.py:120-150) ----------------------------
    for backbone in ("gcn", "gat"):
        torch.manual_seed(51)
        nest = ref.gnn.GNN(in_dim=24, hid_dim=16, num_classes=4, num_layers=2, dropout=0.0, gnn=backbone, device='cpu')
        nest.gnn = nest.init_model()
        with torch.no_grad():
            for p in nest.gnn.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.1, 0.1)
        nest.gnn.train()
        loss, s_logits, t_logits = nest.forward_model(src, tgt)
        nest.gnn.zero_grad()
        loss.backward()
        out["gnn_" + backbone] = {
            "source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
            "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
            "hparams": dict(in_dim=24, hid_dim=16, num_classes=4, num_layers=2, dropout=0.0, gnn=backbone),
            "state": {k: v.clone() for k, v in nest.gnn.state_dict().items()},
            "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
            "target_logits": t_logits.detach().clone(),
            "grads": {k: p.grad.clone() for k, p in nest.gnn.named_parameters() if p.grad is not None}}

    # ---- TDSS (pygda/models/tdss.py): TwoHopNeighbor :21-90, smoothness :314-383, ---------------
    # ---- compute_laplacian_loss :385-449, forward_model :241-312 ------------------------------
    tsrc = small_graph(60, 200, 24, 4, seed=11)
    ttgt = small_graph(50, 150, 24, 4, seed=12)
    tblob = {"source": {"x": tsrc.x, "edge_index": tsrc.edge_index, "y": tsrc.y},
             "target": {"x": ttgt.x, "edge_index": ttgt.edge_index, "y": ttgt.y}}
    hop_in = ref.tdss.Data(edge_index=ttgt.edge_index, edge_attr=None)
    hop_in.num_nodes = 50
    tblob["two_hop"] = ref.tdss.TwoHopNeighbor()(hop_in).edge_index.clone()
    smooth = {}
    for kk in (1, 2, 3):
        e = ref.tdss.TDSS(in_dim=24, hid_dim=16, num_classes=4, smooth_mode='K-hop', k=kk, device='cpu')
        ei_s, _ = e.smoothness(ttgt.edge_index, None, 50)
        smooth[kk] = ei_s.clone()
    tblob["smooth_khop"] = smooth
    torch.manual_seed(31)
    e = ref.tdss.TDSS(in_dim=24, hid_dim=16, num_classes=4, smooth_mode='RW', rw_len=4, device='cpu')
    torch.manual_seed(32)
    walk = __import__("oracle.pyg_ops", fromlist=["random_walk"]).random_walk(
        ttgt.edge_index[0], ttgt.edge_index[1], torch.arange(50), 4)
    torch.manual_seed(32)
    ei_rw, _ = e.smoothness(ttgt.edge_index, None, 50)
    tblob["smooth_rw"] = {"walk": walk, "edge_index": ei_rw.clone(), "seed": 32}
    feats = torch.randn(50, 16, requires_grad=True)
    lap = e.compute_laplacian_loss(feats, smooth[2])
    lap.backward()
    tblob["laplacian"] = {"features": feats.detach().clone(), "edge_index": smooth[2], "loss": lap.detach().clone(),
                          "grad": feats.grad.clone()}
    torch.manual_seed(41)
    hp = dict(in_dim=24, hid_dim=16, num_classes=4, mode='node', smooth_mode='K-hop', num_layers=2, dropout=0.0,
              s_pnums=0, t_pnums=3, k=2, alpha=0.5, beta=0.05)
    test_ = ref.tdss.TDSS(device='cpu', **hp)
    test_.a2gnn = test_.init_model()
    with torch.no_grad():
        for p in test_.a2gnn.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    test_.a2gnn.train()
    ttgt.edge_index_smooth = smooth[2]
    state = {k: v.clone() for k, v in test_.a2gnn.state_dict().items()}
    torch.manual_seed(78)
    loss, s_logits, t_logits = test_.forward_model(tsrc, ttgt, 0.3)
    test_.a2gnn.zero_grad()
    loss.backward()
    torch.manual_seed(78)
    s_idx = torch.randint(60, (5, 1000))
    t_idx = torch.randint(50, (5, 1000))
    tblob.update({"hparams": hp, "alpha_grl": 0.3, "seed": 78, "state": state, "source_idx": s_idx, "target_idx": t_idx,
                  "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
                  "target_logits": t_logits.detach().clone(),
                  "grads": {k: p.grad.clone() for k, p in test_.a2gnn.named_parameters()}})
    out["tdss"] = tblob

    only = sys.argv[1:]
    for name, blob in out.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        torch.save(blob, os.path.join(HERE, name + ".pt"))
        print("wrote", name + ".pt", os.path.getsize(os.path.join(HERE, name + ".pt")), "bytes")


if __name__ == "__main__":
    main()
