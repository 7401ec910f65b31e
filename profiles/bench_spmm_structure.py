#!/usr/bin/env python
"""Which structural feature of the real graph costs the production aggregation kernels their distance to the regular
gather probe (profiles/probes/gather_probe6: 41 us for 1.2M gathers at 64 warps/SM)?  The PRODUCTION kernels
(factored k_spmm_unw via GDA_SPMM_UNW, weighted k_spmm_tasks with GDA_SPMM_UNW=0) are timed on synthetic DIRECTED
graphs whose row (in-degree) structure is controlled:
  regular12   every row 11 uniform-random in-edges + self loop  (the probe's mask 0)
  len_cap64   Chung-Lu row lengths (as config 2), capped at 63 + loop: no long-row segments
  len_cap64p4 the same, every row length rounded up to a multiple of 4 (no tail batches)
  len_real    Chung-Lu row lengths uncapped (hub rows -> 64-nnz segments + ordered reduction), uniform columns
  real        the config-2 target graph itself (power-law columns too)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                                    # noqa: E402
from pygda_b200.graph import Graph                            # noqa: E402
from pygda_b200.synthetic import powerlaw_edge_index          # noqa: E402

n, h, k = 100_000, 128, 10
gen = torch.Generator().manual_seed(0)
real = powerlaw_edge_index(n, 1_000_000, seed=2, offset=48.0)
deg = torch.bincount(real[1], minlength=n)                    # in-degrees of the real graph (without the loop)


def directed(lengths):
    """edge_index with lengths[r] uniform-random in-edges (src != r) for row r."""
    dst = torch.repeat_interleave(torch.arange(n), lengths)
    src = torch.randint(n - 1, (dst.numel(),), generator=gen)
    src = src + (src >= dst).long()                           # never a self loop: the builder adds exactly one
    return torch.stack([src, dst])


cases = {
    "regular12": directed(torch.full((n,), 11)),
    "len_cap64": directed(deg.clamp(max=63)),
    "len_cap64p4": directed(((deg.clamp(max=63) + 1 + 3) // 4 * 4 - 1)),
    "len_real": directed(deg),
    "real": real,
}
only = os.environ.get("CASES")
for name, ei in cases.items():
    if only and name not in only.split(","):
        continue
    g = Graph(ei.cuda(), n)
    x = torch.randn(n, h, device="cuda")
    for _ in range(3):
        ops.spmm_k(g, x, k)
    ops.PROFILE = []
    for _ in range(10):
        ops.spmm_k(g, x, k)
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    us = sum(a.elapsed_time(b) for a, b, _ in recs) / len(recs) * 1e3
    print(f"GDA_SPMM_UNW={os.environ.get('GDA_SPMM_UNW', '12')} GDA_SEG={os.environ.get('GDA_SEG', '64')} {name:12s} kind={recs[0][2][4]:11s} nnz={g.nnz:8d} "
          f"long_rows={g.num_long_rows:5d}: {us:6.1f} us/launch, {us * 1e3 / g.nnz:.4f} ns/nnz, "
          f"gather={4 * g.nnz * h / us / 1e3:8.1f} GB/s")
