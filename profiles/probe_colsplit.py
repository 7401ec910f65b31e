#!/usr/bin/env python
"""Column-split probe: A_hat^k Z is independent per feature column, so the k-step chain can be run on column
slices whose working set (slice of X + slice of Y + indices) stays L2-resident across the chain.
Times k = 10 chained steps at full width (10 launches, H = 128) against 2 x 10 launches on 64-column halves
and 4 x 10 on 32-column quarters (row pitch 128 floats in all cases), same graph as bench_spmm.py."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                                    # noqa: E402
from pygda_b200._lib import gda                               # noqa: E402
from pygda_b200.graph import Graph                            # noqa: E402
from pygda_b200.synthetic import powerlaw_edge_index          # noqa: E402

n, e, H, K = 100_000, 1_000_000, 128, 10
ei = powerlaw_edge_index(n, e, seed=2, offset=48.0).cuda()
g = Graph(ei, n)
x = torch.randn(n, H, device="cuda")
bufs = [torch.empty_like(x), torch.empty_like(x)]
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def chain(parts):
    w = H // parts
    ws = g.workspace(False, w)
    for p in range(parts):
        off = p * w * 4
        src = x.data_ptr() + off
        for i in range(K):
            dst = bufs[i & 1].data_ptr() + off
            gda.spmm_f32(g.handle, 0, C.c_void_p(src), H, C.c_void_p(dst), H, w, None, 0, 0.0, 0, None,
                         C.c_void_p(ws.data_ptr()), ws.numel(), st())
            src = dst
    return bufs[(K - 1) & 1]


ref = None
for parts in (1, 2, 4, 1, 2, 4):
    for _ in range(2):
        out = chain(parts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = chain(parts)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    if ref is None:
        ref = out.clone()
    same = torch.equal(out, ref)
    print(f"parts={parts} width={H // parts:3d}: {us:8.1f} us per A^{K} chain = {us / K:6.1f} us per full-width step  "
          f"bit-identical to full width: {same}")
