"""Layer-1 products at the BASELINE config-2 shape (100 k x 6775, 7 % non-zero, H = 128): tile-packed sparse-X kernels
(csrc/gemm_xt.cu) next to the dense split-bf16 tcgen05 kernel (csrc/gemm_tc.cu).  CUDA events, L2 flushed by the
operands themselves (2.7 GB dense / 0.24 GB packed per pass, output 51 MB)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops                      # noqa: E402
from pygda_b200.data import PackedTiles         # noqa: E402

rows, cols, h = int(os.environ.get("ROWS", 100000)), int(os.environ.get("COLS", 6775)), int(os.environ.get("HID", 128))
dens = float(os.environ.get("DENSITY", 0.07))
torch.manual_seed(0)
x = torch.empty(rows, cols, device="cuda")
for s in range(0, rows, 10000):
    blk = torch.randn(min(10000, rows - s), cols, device="cuda")
    x[s:s + blk.size(0)] = torch.relu(blk) * (torch.rand_like(blk) < 2 * dens)
w = torch.randn(h, cols, device="cuda")
gy = torch.randn(rows, h, device="cuda")
t = PackedTiles(x, pin=False).view()
xs, wsp = ops.Split(x), ops.Split(w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return {"us_median": ts[len(ts) // 2], "us_min": ts[0]}


out = {"shape": [rows, cols, h], "nnz": int(t.vals.numel()), "density": t.vals.numel() / x.numel(),
       "packed_bytes": t.nbytes, "dense_split_bytes": xs.nbytes}
out["xt_fwd"] = timed(lambda: ops.xt_fwd(t, w, False))
out["dense_fwd"] = timed(lambda: ops.gemm_split(xs, wsp, False, True, rows, h, cols))
out["xt_dw"] = timed(lambda: ops.xt_dw(t, gy, False))
gs = ops.Split(gy)
out["dense_dw"] = timed(lambda: ops.gemm_split(gs, xs, True, False, h, cols, rows))
a, _ = ops.xt_fwd(t, w, False)
b = ops.gemm_split(xs, wsp, False, True, rows, h, cols)
out["fwd_identical"] = bool(torch.equal(a, b))
c, _ = ops.xt_dw(t, gy, False)
d = ops.gemm_split(gs, xs, True, False, h, cols, rows)
out["dw_rel_diff"] = float((c - d).abs().max() / d.abs().max())
flops = 3 * 2.0 * rows * cols * h
for k in ("xt_fwd", "dense_fwd", "xt_dw", "dense_dw"):
    out[k]["tensor_TFLOPs"] = flops / out[k]["us_median"] / 1e6
print(json.dumps(out))
