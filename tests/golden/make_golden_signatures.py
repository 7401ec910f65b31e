"""Call signatures of the reference's hot-path classes, read from the reference's own source files:

    python tests/golden/make_golden_signatures.py     # build container only (needs /root/reference)

Writes tests/golden/signatures.json: {"models.A2GNN.__init__": [[name, kind, default-repr | null], ...], ...}.
The drop-in test (tests/test_oracle_golden.py) requires every pygda_b200 signature to START with the reference's
parameters (same names, order, kinds and defaults); extra trailing parameters must be optional."""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402


def sig(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = None if p.default is inspect.Parameter.empty else (p.default.__name__ if callable(p.default) else repr(p.default))
        out.append([p.name, p.kind.name, d])
    return out


def main():
    ref = load_reference()
    targets = {
        "models.BaseGDA": (ref.base.BaseGDA, ["__init__"]),
        "models.A2GNN": (ref.a2gnn.A2GNN, ["__init__", "forward_model", "fit", "predict", "init_model"]),
        "models.UDAGCN": (ref.udagcn.UDAGCN, ["__init__", "forward_model", "fit", "predict", "init_model"]),
        "models.GRADE": (ref.grade.GRADE, ["__init__", "forward_model", "fit", "predict", "init_model"]),
        "models.AdaGCN": (ref.adagcn.AdaGCN, ["__init__", "forward_model", "fit", "predict", "init_model", "gradient_penalty"]),
        "models.GNN": (ref.gnn.GNN, ["__init__", "forward_model", "fit", "predict", "init_model"]),
        "models.TDSS": (ref.tdss.TDSS, ["__init__", "forward_model", "fit", "predict", "smoothness", "compute_laplacian_loss"]),
        "models.DGSDA": (ref.dgsda.DGSDA, ["__init__", "forward_model", "fit", "predict", "entropy_minimization_loss"]),
        "models.StruRW": (ref.strurw.StruRW, ["__init__", "forward_model", "forward_model_mixup", "fit", "predict", "init_model",
                                              "cal_reweight", "cal_edge_prob_sep", "cal_str_dif_rel", "cal_str_diff_ratio",
                                              "shuffle_data", "id_node"]),
        "nn.GCN_reweight": (ref.reweight_gnn.GCN_reweight, ["__init__", "forward"]),
        "nn.GS_reweight": (ref.reweight_gnn.GS_reweight, ["__init__", "forward"]),
        "nn.ReweightGNN": (ref.reweight_gnn.ReweightGNN, ["__init__", "forward"]),
        "nn.MixUpGCNConv": (ref.mixup_gcnconv.MixUpGCNConv, ["__init__", "forward"]),
        "nn.MixupBase": (ref.mixup_base.MixupBase, ["__init__", "forward", "feat_bottleneck", "feat_classifier"]),
        "nn.PropGCNConv": (ref.prop_gcn_conv.PropGCNConv, ["__init__", "forward"]),
        "nn.CachedGCNConv": (ref.cached_gcn_conv.CachedGCNConv, ["__init__", "forward", "norm"]),
        "nn.PPMIConv": (ref.ppmi_conv.PPMIConv, ["__init__", "norm"]),
        "nn.A2GNNBase": (ref.a2gnn_base.A2GNNBase, ["__init__", "forward", "feat_bottleneck", "feat_classifier", "domain_classifier"]),
        "nn.UDAGCNBase": (ref.udagcn_base.UDAGCNBase, ["__init__", "encode", "gcn_encode", "ppmi_encode"]),
        "nn.GRADEBase": (ref.grade_base.GRADEBase, ["__init__", "forward", "feat_bottleneck", "feat_classifier"]),
        "nn.AdaGCNBase": (ref.adagcn_base.AdaGCNBase, ["__init__", "forward"]),
        "nn.GNNBase": (ref.gnn_base.GNNBase, ["__init__", "forward"]),
        "nn.BernProp": (ref.dgsda_base.BernProp, ["__init__", "forward"]),
        "nn.DGSDABase": (ref.dgsda_base.DGSDABase, ["__init__", "forward", "get_props"]),
        "nn.Attention": (ref.attention.Attention, ["__init__", "forward"]),
    }
    out = {}
    for name, (cls, methods) in targets.items():
        for m in methods:
            out[f"{name}.{m}"] = sig(getattr(cls, m))
    out["utils.MMD"] = sig(ref.mmd.MMD)
    out["utils.logger"] = sig(ref.utility.logger)
    out["nn.gcn_norm"] = sig(ref.prop_gcn_conv.gcn_norm)
    json.dump(out, open(os.path.join(HERE, "signatures.json"), "w"), indent=0)
    print("wrote signatures.json", len(out), "signatures")


if __name__ == "__main__":
    main()
