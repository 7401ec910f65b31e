"""Golden values of the reference's own host-side metrics (pygda/metrics/metrics.py):

    python tests/golden/make_golden_metrics.py      # build container only (needs /root/reference)

Writes tests/golden/metrics.json: seeded inputs (as lists) and the value of every exported metric."""
import importlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402


def main():
    load_reference()
    RM = importlib.import_module("pygda.metrics.metrics")
    g = torch.Generator().manual_seed(0)
    out = []
    for n, p in ((50, 0.3), (200, 0.1), (64, 0.6)):
        label = (torch.rand(n, generator=g) < p).long()
        score = torch.rand(n, generator=g)
        for sc, tag in ((score, "score"), (-score, "negated")):
            out.append({"label": label.tolist(), "score": sc.tolist(), "tag": tag,
                        "eval_roc_auc": float(RM.eval_roc_auc(label, sc)),
                        "eval_recall_at_k": float(RM.eval_recall_at_k(label, sc)),
                        "eval_precision_at_k": float(RM.eval_precision_at_k(label, sc, 7)),
                        "eval_average_precision": float(RM.eval_average_precision(label, sc))})
        y, pred = torch.randint(4, (n,), generator=g), torch.randint(4, (n,), generator=g)
        out.append({"label": y.tolist(), "pred": pred.tolist(), "tag": "multiclass",
                    "eval_micro_f1": float(RM.eval_micro_f1(y, pred)), "eval_macro_f1": float(RM.eval_macro_f1(y, pred))})
    json.dump(out, open(os.path.join(HERE, "metrics.json"), "w"))
    print("wrote metrics.json", len(out))


if __name__ == "__main__":
    main()
