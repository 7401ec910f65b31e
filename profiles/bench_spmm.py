#!/usr/bin/env python
"""Micro-benchmark of the aggregation kernel on the config-2 target graph (100k nodes, 1.1M nnz):
CUDA-event time per launch in a chain of 50 launches (L2-warm, like the k-step propagation),
B_alg GB/s, and a correctness check against torch index_add on the GPU."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                                    # noqa: E402
from pygda_b200.graph import Graph                            # noqa: E402
from pygda_b200.synthetic import powerlaw_edge_index          # noqa: E402

n = int(os.environ.get("N", 100_000)); e = int(os.environ.get("E", 1_000_000))
cases = ((128, torch.float32), (5, torch.float32), (256, torch.bfloat16), (64, torch.float32))
if os.environ.get("WIDE_ONLY"):                                # the one-16-byte-slice-per-lane widths only
    cases = ((128, torch.float32), (256, torch.bfloat16))
print("GDA_SPMM_TASKS =", os.environ.get("GDA_SPMM_TASKS", "(default 4)"), " SAME_X =", os.environ.get("SAME_X", "(ping-pong)"))
for h, dt in cases:
    ei = powerlaw_edge_index(n, e, seed=2, offset=48.0).cuda()
    g = Graph(ei, n)
    x = torch.randn(n, h, device="cuda").to(dt)
    if os.environ.get("X_ZEROS"):                              # data-dependence check (L2 / fabric behaviour)
        x.zero_()
    y = ops.spmm(g, x)
    cei, cw = g.coo()
    ref = torch.zeros(n, h, device="cuda").index_add_(0, cei[1], x.float()[cei[0]] * cw[:, None])
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    bufs = [x, torch.empty_like(x)]
    for _ in range(5):
        ops.spmm(g, bufs[0], out=bufs[1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    same_x = bool(os.environ.get("SAME_X"))                     # A/B for round 2: X read-only (never re-written) vs the
    e0.record()                                                # ping-pong chain, whose input was just written by the
    for i in range(reps):                                      # previous launch (home-L2 / cross-die fabric hypothesis)
        if same_x:
            ops.spmm(g, bufs[0], out=bufs[1])
        else:
            ops.spmm(g, bufs[i & 1], out=bufs[(i + 1) & 1])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    s = x.element_size()
    balg = 4 * (n + 1) + 8 * g.nnz + 2 * s * n * h
    print(f"H={h:4d} {str(dt):15s} nnz={g.nnz} long_rows={g.num_long_rows}  {us:8.1f} us/launch  "
          f"B_alg={balg / 1e6:7.1f} MB -> {balg / us / 1e3:7.1f} GB/s   gather={s * g.nnz * h / us / 1e3:7.1f} GB/s  relerr={err:.1e}")
