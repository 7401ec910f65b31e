"""Attributes the reference's estimator constructors set (names and simple values), from the reference's own files:

    python tests/golden/make_golden_attrs.py      # build container only (needs /root/reference)

Writes tests/golden/estimator_attrs.json: {"A2GNN": {"kwargs": {...}, "attrs": {name: value}}, ...} for values of
type int / float / str / bool / None / list of those."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

SIMPLE = (int, float, str, bool, type(None))


def simple(v):
    return isinstance(v, SIMPLE) or (isinstance(v, list) and all(isinstance(x, SIMPLE) for x in v))


def main():
    ref = load_reference()
    classes = {"A2GNN": ref.a2gnn.A2GNN, "UDAGCN": ref.udagcn.UDAGCN, "GRADE": ref.grade.GRADE, "AdaGCN": ref.adagcn.AdaGCN,
               "GNN": ref.gnn.GNN, "TDSS": ref.tdss.TDSS, "DGSDA": ref.dgsda.DGSDA, "StruRW": ref.strurw.StruRW}
    out = {}
    for name, cls in classes.items():
        for tag, kw in (("defaults", dict(in_dim=12, hid_dim=8, num_classes=3, device="cpu")),
                        ("custom", dict(in_dim=12, hid_dim=8, num_classes=3, device="cpu", num_layers=2, lr=0.02,
                                        weight_decay=0.003, epoch=7, batch_size=0, num_neigh=[5, 3], verbose=1, dropout=0.25))):
            est = cls(**kw)
            out[f"{name}/{tag}"] = {"kwargs": kw, "attrs": {k: v for k, v in vars(est).items() if simple(v)}}
    json.dump(out, open(os.path.join(HERE, "estimator_attrs.json"), "w"), indent=0, sort_keys=True)
    print("wrote estimator_attrs.json", len(out))


if __name__ == "__main__":
    main()
