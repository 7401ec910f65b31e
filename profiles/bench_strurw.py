#!/usr/bin/env python
"""StruRW on one B200 at the config-2 shape: per domain 100k nodes / 1M directed edges / 6775 features /
5 classes, hid 128, 2 layers, GS backbone with 'mean' pooling and the 'erm' objective (the defaults of
benchmark/node/strurw.py:37-46), lr 0.003 / wd 0.01 / dropout 0.2 / lamb 0.8 (benchmark/node/run_citation.sh:6).
CUDA-event time per training step with and without a re-weighting in the step, the edge re-weighting alone
(reference: dense N x N adjacencies on the host, 40 GB each at this size), per-launch aggregation time and B_alg GB/s on
the folded re-weighted CSR (isolated nodes get an explicit zero-weight entry, so the lean kernels apply).  One JSON object per line on stdout."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                                    # noqa: E402
from pygda_b200.models import StruRW                          # noqa: E402
from pygda_b200.nn.reweight_gnn import message_graph          # noqa: E402
from pygda_b200.optim import Adam                             # noqa: E402
from pygda_b200.synthetic import domain_pair                  # noqa: E402

N = int(os.environ.get("N", 100_000)); E = int(os.environ.get("E", 1_000_000))
F_, H, C = int(os.environ.get("F", 6775)), 128, 5
dev = torch.device("cuda:0")
peak = 6554.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


src, tgt = domain_pair(N, E, F_, C, seed=0, device=dev)
for gnn in ("GS", "GCN"):
    torch.manual_seed(0)
    est = StruRW(in_dim=F_, hid_dim=H, num_classes=C, num_layers=2, cls_dim=128, cls_layers=2, dropout=0.2, gnn=gnn,
                 pooling="mean", lamb=0.8, mode="erm", ew_start=1, ew_freq=1, lr=0.003, weight_decay=0.01, epoch=200,
                 device=str(dev), verbose=0)
    est.gnn = est.init_model()
    opt = Adam(est.gnn.parameters(), lr=0.003, weight_decay=0.01)
    s, t = est._to_device(src), est._to_device(tgt)
    est.reweight = False
    plain = lambda: est.train_step(s, t, 5, opt)                    # noqa: E731
    for _ in range(3):
        loss = plain()[0]
    ms_plain = timed(plain, 10)
    est.reweight = True                                             # ew_freq = 1: every step re-weights the source edges
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):                 # 'edge reweight...' once per step
        for _ in range(2):
            plain()
        ms_rw = timed(plain, 10)
        pred = torch.randint(C, (N,), device=dev)
        ms_cal = timed(lambda: est.cal_reweight(s, t, pred), 10)
    est.reweight = False
    ops.PROFILE = []
    for _ in range(3):
        plain()
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    tm = [a.elapsed_time(b) for a, b, meta in recs if meta[:2] == (N, H)]
    g = message_graph(t.edge_index, t.edge_weight, 0.8, N, gnn == "GCN", "mean")
    b_alg = 4 * (N + 1) + 8 * g.nnz + 2 * 4 * N * H
    us = statistics.mean(tm) * 1e3 if tm else float("nan")
    print(json.dumps({"what": "StruRW training step (erm, mean pooling)", "backbone": gnn, "nodes": N, "edges": E,
                      "feat": F_, "hid": H, "ms_per_step": ms_plain, "epochs_per_s": 1e3 / ms_plain,
                      "ms_per_step_with_reweighting": ms_rw, "ms_cal_reweight": ms_cal, "loss": float(loss.detach()),
                      "aggregation": {"launches_per_step": len(tm) / 3, "us_per_launch": us, "nnz": g.nnz,
                                      "alg_bytes_per_launch": b_alg, "achieved_GBps": b_alg / us / 1e3,
                                      "frac_of_hbm_peak": b_alg / us / 1e3 / peak,
                                      "gather_GBps": 4 * g.nnz * H / us / 1e3},
                      "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    del est, opt
    torch.cuda.empty_cache()
