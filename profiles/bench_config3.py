#!/usr/bin/env python
"""BASELINE.json config 3 on one B200: UDAGCN (GRL + discriminator, ppmi=False) on synthetic citation-shaped
graphs, per domain 1M nodes / 10M directed edges / 512 features / 5 classes, hid 256, 2 layers
(benchmark/node/run_citation.sh:90 hyper-parameters: lr 1e-4, wd 1e-3) -- fp32 features vs the bf16 feature path
(`UDAGCN(feature_dtype=torch.bfloat16)`), CUDA-event time per training step, per-launch aggregation time and
B_alg GB/s, plus the GPU PPMI graph build on the same graph (the reference's pure-Python
PPMIConv.norm would take hours here) and an aggregation size sweep (is the working set L2-resident?).
One JSON object per line on stdout."""
import itertools
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                                    # noqa: E402
from pygda_b200.graph import Graph, clear_graph_cache         # noqa: E402
from pygda_b200.models import UDAGCN                          # noqa: E402
from pygda_b200.optim import Adam                             # noqa: E402
from pygda_b200.synthetic import domain_pair, powerlaw_edge_index   # noqa: E402

N = int(os.environ.get("N", 1_000_000)); E = int(os.environ.get("E", 10_000_000))
F_, H, C = 512, 256, 5
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


if os.environ.get("PROFILE_ONE"):
    # ncu --profile-from-start off: one training step of one variant between cudaProfilerStart/Stop, then exit
    src, tgt = domain_pair(N, E, F_, C, seed=0, device=dev)
    dt = torch.bfloat16 if os.environ["PROFILE_ONE"] == "bf16" else None
    torch.manual_seed(0)
    est = UDAGCN(in_dim=F_, hid_dim=H, num_classes=C, num_layers=2, ppmi=False, lr=1e-4, weight_decay=1e-3,
                 epoch=400, device=str(dev), verbose=0, feature_dtype=dt)
    est.udagcn = est.init_model()
    opt = Adam(itertools.chain(*[m.parameters() for m in est.udagcn.models]), lr=1e-4, weight_decay=1e-3)
    for _ in range(2):
        est.train_step(src, tgt, 0.05, 10, opt)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    est.train_step(src, tgt, 0.05, 10, opt)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

if not os.environ.get("SKIP_STEP"):
    src, tgt = domain_pair(N, E, F_, C, seed=0, device=dev)
    for name, dt in (("fp32", None), ("bf16", torch.bfloat16)):
        torch.manual_seed(0)
        est = UDAGCN(in_dim=F_, hid_dim=H, num_classes=C, num_layers=2, ppmi=False, lr=1e-4, weight_decay=1e-3,
                     epoch=400, device=str(dev), verbose=0, feature_dtype=dt)
        est.udagcn = est.init_model()
        opt = Adam(itertools.chain(*[m.parameters() for m in est.udagcn.models]), lr=1e-4, weight_decay=1e-3)
        step = lambda: est.train_step(src, tgt, 0.05, 10, opt)          # noqa: E731
        for _ in range(3):
            loss = step()[0]
        ms = timed(step, 10)
        ops.PROFILE = []
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        recs, ops.PROFILE = ops.PROFILE, None
        tm = [a.elapsed_time(b) for a, b, meta in recs if meta[:2] == (N, H)]
        g = Graph(tgt.edge_index, N, None, 1 | 2)               # SELF_LOOPS | NORM_SYM_ROW, as CachedGCNConv builds it
        s = 2 if dt is not None else 4
        b_alg = 4 * (N + 1) + 8 * g.nnz + 2 * s * N * H
        us = statistics.mean(tm) * 1e3
        print(json.dumps({"what": "UDAGCN config 3 training step", "features": name, "nodes": N, "edges": E, "feat": F_,
                          "hid": H, "ms_per_step": ms, "epochs_per_s": 1e3 / ms, "loss": float(loss),
                          "aggregation": {"launches_per_step": len(tm) / 3, "us_per_launch": us, "nnz": g.nnz,
                                          "alg_bytes_per_launch": b_alg, "achieved_GBps": b_alg / us / 1e3,
                                          "frac_of_hbm_peak": b_alg / us / 1e3 / peak,
                                          "gather_GBps": s * g.nnz * H / us / 1e3},
                          "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
        del est, opt, g
        clear_graph_cache()
        torch.cuda.empty_cache()

    # ---- PPMI graph of the target graph on the GPU (pygda/nn/ppmi_conv.py:98-172; path_len 10 as UDAGCN uses) ----
    from pygda_b200.ppmi import ppmi_edges
    ppmi_edges(tgt.edge_index[:, :1000].contiguous(), N, 10, seed=1)           # warm the context
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pe, pw = ppmi_edges(tgt.edge_index, N, 10, seed=1)
    torch.cuda.synchronize()
    t_ppmi = time.perf_counter() - t0
    print(json.dumps({"what": "PPMI graph build on the GPU (40 rounds, path_len 10)", "nodes": N, "edges": E,
                      "seconds": t_ppmi, "ppmi_edges": int(pe.size(1)), "walk_steps": 40 * N * 5.5,
                      "nonzero_scores": int((pw > 0).sum())}), flush=True)
    del src, tgt, pe, pw
    torch.cuda.empty_cache()

# ---- aggregation size sweep at the config-2 shape family (mean degree 10, H = 128 fp32) ----
for n in (() if os.environ.get("SKIP_SWEEP") else (12_500, 25_000, 50_000, 100_000, 200_000, 400_000, 800_000)):
    ei = powerlaw_edge_index(n, 10 * n, seed=2, offset=48.0).cuda()
    g = Graph(ei, n)
    x = torch.randn(n, 128, device=dev)
    bufs = [x, torch.empty_like(x)]
    it = itertools.count()

    def one():
        i = next(it)
        ops.spmm(g, bufs[i & 1], out=bufs[(i + 1) & 1])
    for _ in range(6):
        one()
    us = timed(one, 50) * 1e3
    ws = (2 * 4 * n * 128 + 8 * g.nnz) / 1e6
    print(json.dumps({"what": "aggregation size sweep (H=128 fp32, chained launches)", "nodes": n, "nnz": g.nnz,
                      "working_set_MB": ws, "us_per_launch": us, "ns_per_nnz": us * 1e3 / g.nnz,
                      "gather_TBps": 4 * g.nnz * 128 / us / 1e6}), flush=True)
    del g, x, bufs
