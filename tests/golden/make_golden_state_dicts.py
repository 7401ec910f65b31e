"""state_dict layouts (parameter / buffer names and shapes) of the reference's hot-path modules, from its own files:

    python tests/golden/make_golden_state_dicts.py      # build container only (needs /root/reference)

Writes tests/golden/state_dicts.json: {"A2GNNBase/node": {"kwargs": {...}, "state": {key: [shape]}}, ...}."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402


def main():
    ref = load_reference()
    cases = {
        "A2GNNBase/node": (ref.a2gnn_base.A2GNNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=3, adv=True, mode="node")),
        "A2GNNBase/graph": (ref.a2gnn_base.A2GNNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, adv=False, mode="graph")),
        "UDAGCNBase/ppmi": (ref.udagcn_base.UDAGCNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, ppmi=True, adv_dim=10)),
        "UDAGCNBase/plain": (ref.udagcn_base.UDAGCNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=3, ppmi=False, adv_dim=10)),
        "GRADEBase/JS": (ref.grade_base.GRADEBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, disc="JS")),
        "GRADEBase/C": (ref.grade_base.GRADEBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, disc="C")),
        "AdaGCNBase/gcn": (ref.adagcn_base.AdaGCNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2)),
        "AdaGCNBase/ppmi": (ref.adagcn_base.AdaGCNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, gnn_type="ppmi")),
        "GNNBase/gcn": (ref.gnn_base.GNNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, gnn="gcn")),
        "GNNBase/gat": (ref.gnn_base.GNNBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=2, gnn="gat")),
        "DGSDABase": (ref.dgsda_base.DGSDABase, dict(features=12, hidden=8, classes=3, K=5)),
        "PropGCNConv": (ref.prop_gcn_conv.PropGCNConv, dict(in_channels=12, out_channels=8)),
        "CachedGCNConv": (ref.cached_gcn_conv.CachedGCNConv, dict(in_channels=12, out_channels=8)),
        "PPMIConv": (ref.ppmi_conv.PPMIConv, dict(in_channels=12, out_channels=8, path_len=7)),
        "Attention": (ref.attention.Attention, dict(in_channels=8)),
        "ReweightGNN/GS": (ref.reweight_gnn.ReweightGNN, dict(input_dim=12, gnn_dim=8, output_dim=3, cls_dim=6, gnn_layers=3, cls_layers=2, backbone="GS")),
        "ReweightGNN/GCN": (ref.reweight_gnn.ReweightGNN, dict(input_dim=12, gnn_dim=8, output_dim=3, cls_dim=6, gnn_layers=2, cls_layers=3, backbone="GCN", pooling="add")),
        "GCN_reweight": (ref.reweight_gnn.GCN_reweight, dict(in_channels=12, out_channels=8, aggr="mean")),
        "GS_reweight": (ref.reweight_gnn.GS_reweight, dict(in_channels=12, out_channels=8, reducer="mean")),
        "MixUpGCNConv": (ref.mixup_gcnconv.MixUpGCNConv, dict(in_channels=12, out_channels=8)),
        "MixupBase": (ref.mixup_base.MixupBase, dict(in_dim=12, hid_dim=8, num_classes=3, num_layers=3)),
    }
    out = {}
    for name, (cls, kw) in cases.items():
        torch.manual_seed(0)
        m = cls(**kw)
        out[name] = {"kwargs": kw, "state": {k: list(v.shape) for k, v in m.state_dict().items()},
                     "trainable": sorted(k for k, p in m.named_parameters() if p.requires_grad)}
    json.dump(out, open(os.path.join(HERE, "state_dicts.json"), "w"), indent=0, sort_keys=True)
    print("wrote state_dicts.json", len(out))


if __name__ == "__main__":
    main()
