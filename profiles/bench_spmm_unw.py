#!/usr/bin/env python
"""A_hat^10 on the config-2 target graph (100k nodes / 1.1M nnz, H = 128 fp32): the factored unit-weight chain
(k_spmm_unw, GDA_SPMM_UNW = 12 default / 16 / 0 = weighted k_spmm_tasks) timed per step with CUDA events, checked
against the weighted chain.  N / E / NB environment variables change the size / stack two matrices."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                                    # noqa: E402
from pygda_b200.graph import Graph                            # noqa: E402
from pygda_b200.synthetic import powerlaw_edge_index          # noqa: E402

n = int(os.environ.get("N", 100_000)); e = int(os.environ.get("E", 1_000_000)); nb = int(os.environ.get("NB", 1))
k, h = 10, 128
ei = powerlaw_edge_index(n, e, seed=2, offset=48.0).cuda()
g = Graph(ei, n)
x = torch.randn(nb * n, h, device="cuda")
mode = os.environ.get("GDA_SPMM_UNW", "(default 16)")
fact = ops.unit_weight_chain(g, False, h, nb, k)
y = ops.spmm_k(g, x, k, nb=nb)
w = x
for _ in range(k):
    w = ops.spmm(g, w, nb=nb)
err = float((y - w).abs().max() / w.abs().max())
for _ in range(3):
    ops.spmm_k(g, x, k, nb=nb)
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.spmm_k(g, x, k, nb=nb)
e1.record()
torch.cuda.synchronize()
us_chain = e0.elapsed_time(e1) / reps * 1e3
ops.PROFILE = []
for _ in range(reps):
    ops.spmm_k(g, x, k, nb=nb)
torch.cuda.synchronize()
recs, ops.PROFILE = ops.PROFILE, None
us = sum(a.elapsed_time(b) for a, b, _ in recs) / len(recs) * 1e3
kind = recs[0][2][4]
idx = 4 * (n + 1) + (4 * g.nnz + 4 * n if kind == "unit-weight" else 8 * g.nnz)
balg = idx + nb * 2 * 4 * n * h
print(f"GDA_SPMM_UNW={mode} factored={fact} kind={kind} nb={nb} nnz={g.nnz}: chain of {k} = {us_chain:7.1f} us "
      f"({us_chain / k:6.1f} us/step incl. row scaling), {us:6.1f} us/launch (events), B_alg={balg / 1e6:.1f} MB -> "
      f"{balg / us / 1e3:7.1f} GB/s, gather={nb * 4 * g.nnz * h / us / 1e3:8.1f} GB/s, relerr vs weighted={err:.1e}")
