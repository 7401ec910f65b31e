"""Sustained loops of the layer-1 kernels with NVML clock / power sampling: are they power-limited?"""
import json, os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200 import ops
from pygda_b200.data import PackedTiles
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
rows, cols, hid = 100000, 6775, 128
torch.manual_seed(0)
x = torch.empty(rows, cols, device="cuda")
for s in range(0, rows, 10000):
    blk = torch.randn(min(10000, rows - s), cols, device="cuda")
    x[s:s + blk.size(0)] = torch.relu(blk) * (torch.rand_like(blk) < 0.14)
w = torch.randn(hid, cols, device="cuda"); gy = torch.randn(rows, hid, device="cuda")
t = PackedTiles(x, pin=False).view()
tz = PackedTiles(x, pin=False).view(); tz.vals.zero_()         # same pattern, all values 0.0 (kept: bit pattern irrelevant here)
xs, wsp, gs = ops.Split(x), ops.Split(w), ops.Split(gy)
def run(name, fn, secs=0.6):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    samples, stop = [], [False]
    def samp():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.004)
    th = threading.Thread(target=samp); th.start()
    n = 0; t0 = time.perf_counter(); ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    ev1.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    cl = sorted(c for c, _ in samples); pw = sorted(p for _, p in samples)
    print(json.dumps({"kernel": name, "us_per_launch": ev0.elapsed_time(ev1) * 1e3 / n, "launches": n,
                      "sm_mhz_median": cl[len(cl) // 2], "sm_mhz_min": cl[0], "power_w_median": pw[len(pw) // 2], "power_w_max": pw[-1]}))
run("xt_fwd", lambda: ops.xt_fwd(t, w, False))
run("xt_fwd_zero_values", lambda: ops.xt_fwd(tz, w, False))
run("dense_fwd", lambda: ops.gemm_split(xs, wsp, False, True, rows, hid, cols))
run("xt_dw", lambda: ops.xt_dw(t, gy, False))
run("xt_dw_zero_values", lambda: ops.xt_dw(tz, gy, False))
run("dense_dw", lambda: ops.gemm_split(gs, xs, True, False, hid, cols, rows))
