"""Golden vectors for the re-weighted conv family and StruRW, made by EXECUTING THE REFERENCE'S OWN
FILES (pygda/nn/reweight_gnn.py, pygda/nn/mixup_gcnconv.py, pygda/nn/mixup_base.py, pygda/models/strurw.py):

    python tests/golden/make_golden_strurw.py        # build container only (needs /root/reference)

Writes tests/golden/strurw.pt:
  layers   GS_reweight ('mean' / 'add'), GCN_reweight ('mean' / 'add'), MixUpGCNConv: forward + all gradients on a
           graph with duplicate edges, self loops and nodes without outgoing edges, non-trivial edge weights
  nets     ReweightGNN (GS and GCN backbones, 1- and 3-layer classifier heads) and MixupBase.feat_bottleneck
           (2 and 3 layers, a fixed permutation and lam): outputs + all parameter gradients
  reweight StruRW.cal_reweight: the source edge weights for given target predictions (float32, exact)
  strurw   StruRW.forward_model for mode erm / adv / mmd at an epoch where the edge re-weighting fires, and
           forward_model_mixup with np.random seeded: loss, logits, new edge weights, all gradients
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference, REPO  # noqa: E402

sys.path.insert(0, REPO)
from make_golden import small_graph  # noqa: E402
from oracle.data import Data  # noqa: E402


def ragged_graph(n, e, f, c, seed):
    """small_graph + a handful of nodes whose OUT-edges are removed (empty rows for the target_to_source layers)."""
    g = small_graph(n, e, f, c, seed=seed)
    ei = g.edge_index
    keep = ~((ei[0] == 2) | (ei[0] == 5) | (ei[0] == n - 1))
    return Data(x=g.x, edge_index=ei[:, keep].contiguous(), y=g.y)


def grads_of(mod):
    return {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}


def state_of(mod):
    return {k: v.clone() for k, v in mod.state_dict().items()}


def main():
    ref = load_reference()
    R, MX = ref.reweight_gnn, ref.mixup_gcnconv
    out = {}

    # ---- single layers ---------------------------------------------------------------------
    torch.manual_seed(21)
    g = ragged_graph(60, 220, 10, 3, seed=13)
    E = g.edge_index.size(1)
    ew = torch.rand(E) * 1.5 + 0.25
    lmda = 0.8
    layers = {}
    for name, make in {
        "gs_mean": lambda: R.GS_reweight(10, 7, "mean"),
        "gs_add": lambda: R.GS_reweight(10, 7, "add"),
        "gs_mean_normalized": lambda: R.GS_reweight(10, 7, "mean", normalize_embedding=True),
        "gcn_mean": lambda: R.GCN_reweight(10, 7, "mean"),
        "gcn_add": lambda: R.GCN_reweight(10, 7, "add"),
    }.items():
        layer = make()
        with torch.no_grad():
            for p in layer.parameters():
                if p.dim() == 1:
                    p.copy_(torch.randn_like(p) * 0.1)           # biases away from their zero init
        x = g.x.clone().requires_grad_(True)
        y = layer(x, g.edge_index, ew, lmda)
        go = torch.randn_like(y)
        y.backward(go)
        layers[name] = {"state": state_of(layer), "y": y.detach().clone(), "gout": go, "gx": x.grad.clone(),
                        "grads": grads_of(layer)}
    conv = MX.MixUpGCNConv(10, 7)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(7) * 0.1)
    x = g.x.clone().requires_grad_(True)
    xc = torch.randn(60, 10, requires_grad=True)
    y = conv(x, xc, g.edge_index, ew, lmda)
    go = torch.randn_like(y)
    y.backward(go)
    layers["mixup_conv"] = {"state": state_of(conv), "x_cen": xc.detach().clone(), "y": y.detach().clone(),
                            "gout": go, "gx": x.grad.clone(), "gx_cen": xc.grad.clone(), "grads": grads_of(conv)}
    out["layers"] = {"x": g.x, "edge_index": g.edge_index, "edge_weight": ew, "lmda": lmda, "num_nodes": 60,
                     "cases": layers}

    # ---- networks --------------------------------------------------------------------------
    nets = {}
    data = Data(x=g.x, edge_index=g.edge_index, y=g.y, edge_weight=ew)
    for name, kw in {"gs": dict(backbone="GS", pooling="mean", gnn_layers=3, cls_layers=2),
                     "gcn": dict(backbone="GCN", pooling="mean", gnn_layers=2, cls_layers=3),
                     "gcn_add": dict(backbone="GCN", pooling="add", gnn_layers=2, cls_layers=1),
                     "gs_bn": dict(backbone="GS", pooling="mean", gnn_layers=2, cls_layers=2, bn=True)}.items():
        torch.manual_seed(31)
        hp = dict(input_dim=10, gnn_dim=8, output_dim=3, cls_dim=6, dropout=0.0, rw_lmda=0.6, **kw)
        net = R.ReweightGNN(**hp)
        net.train()
        st = state_of(net)
        feat, logits = net(data, data.x)
        gf, gl = torch.randn_like(feat), torch.randn_like(logits)
        (feat * gf).sum().add((logits * gl).sum()).backward()
        nets[name] = {"hparams": hp, "state": st, "feat": feat.detach().clone(), "logits": logits.detach().clone(),
                      "gfeat": gf, "glogits": gl, "grads": grads_of(net),
                      "state_after": state_of(net)}                      # BatchNorm running stats (bn=True)
    perm = np.random.RandomState(3).permutation(60)
    eib = ragged_graph(60, 220, 10, 3, seed=14).edge_index[:, :E]
    if eib.size(1) < E:                                                  # same edge count as edge_weight
        eib = torch.cat([eib, g.edge_index[:, : E - eib.size(1)]], 1)
    for L in (2, 3):
        torch.manual_seed(37)
        hp = dict(in_dim=10, hid_dim=8, num_classes=3, num_layers=L, dropout=0.0, rw_lmda=0.7)
        net = ref.mixup_base.MixupBase(**hp)
        net.train()
        st = state_of(net)
        feat = net.feat_bottleneck(g.x, g.edge_index, eib, 0.35, perm, ew)
        logits = net.feat_classifier(feat)
        gl = torch.randn_like(logits)
        (logits * gl).sum().backward()
        nets[f"mixup{L}"] = {"hparams": hp, "state": st, "edge_index_b": eib, "lam": 0.35,
                             "perm": torch.from_numpy(perm), "feat": feat.detach().clone(),
                             "logits": logits.detach().clone(), "glogits": gl, "grads": grads_of(net)}
    out["nets"] = nets

    # ---- StruRW ----------------------------------------------------------------------------
    src = ragged_graph(70, 260, 12, 4, seed=5)
    tgt = small_graph(64, 230, 12, 4, seed=6, with_loops=False)

    def fresh(d):
        return Data(x=d.x, edge_index=d.edge_index, y=d.y, edge_weight=torch.ones(d.edge_index.size(1)))

    # cal_reweight alone (pygda/models/strurw.py:472-505): one class never predicted on the target (0/0 -> 1)
    est = ref.strurw.StruRW(in_dim=12, hid_dim=8, num_classes=4, device="cpu")
    s, t = fresh(src), fresh(tgt)
    pred = torch.from_numpy(np.random.RandomState(9).randint(0, 3, size=64)).long()
    est.cal_reweight(s, t, pred)
    out["reweight"] = {"source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
                       "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
                       "target_pred": pred, "edge_weight": s.edge_weight.clone(), "num_classes": 4}

    runs = {}
    for mode, gnn in (("erm", "GS"), ("adv", "GS"), ("mmd", "GCN"), ("mixup", None)):
        hp = dict(in_dim=12, hid_dim=8, num_classes=4, num_layers=2, cls_dim=6, cls_layers=2, dropout=0.0,
                  pooling="mean", reweight=True, pseudo=True, ew_start=2, ew_freq=2, lamb=0.8, mode=mode)
        if gnn is not None:
            hp["gnn"] = gnn
        torch.manual_seed(51)
        est = ref.strurw.StruRW(device="cpu", **hp)
        est.gnn = est.init_model()
        est.gnn.train()
        state = state_of(est.gnn)
        mods = [est.gnn]
        if mode == "adv":
            est.domain_discriminator = torch.nn.Linear(8, 2)
            mods.append(est.domain_discriminator)
            dstate = state_of(est.domain_discriminator)
        s, t = fresh(src), fresh(tgt)
        torch.manual_seed(53)                                # MMD sample indices (pygda/utils/mmd.py:148-149)
        np.random.seed(57)                                   # mixup: beta draw + node shuffle (strurw.py:301,723)
        if mode == "mixup":
            loss, s_logits, t_logits = est.forward_model_mixup(s, t, 1)
        else:
            loss, s_logits, t_logits = est.forward_model(s, t, 0.37, 1)     # (epoch + 1) % ew_freq == 0
        for m in mods:
            m.zero_grad()
        loss.backward()
        runs[mode] = {"hparams": hp, "state": state, "alpha": 0.37, "epoch": 1, "seed": 53, "np_seed": 57,
                      "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
                      "target_logits": t_logits.detach().clone(), "edge_weight": s.edge_weight.clone(),
                      "grads": grads_of(est.gnn)}
        if mode == "adv":
            runs[mode]["disc_state"] = dstate
            runs[mode]["disc_grads"] = grads_of(est.domain_discriminator)
    out["strurw"] = {"source": out["reweight"]["source"], "target": out["reweight"]["target"], "runs": runs}

    torch.save(out, os.path.join(HERE, "strurw.pt"))
    print("wrote strurw.pt", os.path.getsize(os.path.join(HERE, "strurw.pt")), "bytes")


if __name__ == "__main__":
    main()
