#!/usr/bin/env python
"""bench.py -- A2GNN training epochs/sec on the BASELINE.json config-2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (hyper-parameters of
benchmark/node/run_citation.sh:92): synthetic citation-shaped source/target graphs,
per domain 100 000 nodes / 1 000 000 directed edges / 6 775 features / 5 classes, fp32;
A2GNN with 2 PropGCNConv layers, hid 128, s_pnums 0, t_pnums 10, dropout 0.5, MMD weight 10,
Adam lr 0.01 wd 0.005.  Full-batch node mode: one epoch == one optimiser step
(pygda/models/a2gnn.py:300-319); per-epoch sklearn F1 + printing are excluded from both arms.

One JSON line on stdout (rank 0):
  value     epochs/sec with the inputs resident in HBM (CUDA events, max over ranks)
  e2e       the same through the estimator API from pinned HOST buffers: every step copies
            x / edge_index / y of both graphs host->device and reads the loss back.  `Data.pin_memory()`
            keeps a sparse x row-compressed in pinned memory (lossless, rebuilt densely on the GPU);
            e2e.dense_staging is the same measurement with the plain dense pinned copy
  roofline  aggregation kernel (A_hat x, H=128, target graph): algorithmic bytes per launch
            B_alg = 4(N+1) + 8 nnz + 2*4*N*H  divided by the mean per-launch CUDA-event time
            measured in an instrumented repeat of the timed steps; peak = MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restatement of the reference's torch op sequence) on this
            box's host cores, on a bounded sample (see `sample`)
`--impl reference` prints the reference arm alone: the oracle on the host CPU.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(nodes=100_000, edges=1_000_000, feats=6775, classes=5, hid=128, layers=2, s_pnums=0,
           t_pnums=10, dropout=0.5, weight=10, lr=0.01, weight_decay=0.005, epochs=200)
METRIC = "a2gnn_train_epochs_per_sec"
UNIT = "epochs/s"


def workload_name():
    return ("A2GNN synthetic citation-shape %dk nodes / %dM edges / %d feat / %d classes, fp32 "
            "(BASELINE.json configs[1])" % (CFG["nodes"] // 1000, CFG["edges"] // 1_000_000,
                                            CFG["feats"], CFG["classes"]))


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.thread, self.gpu = [], None, None, str(gpu_index)
        self.nvml_rows, self.nvml_thread, self.nvml_stop = [], None, threading.Event()

    def _nvml_pump(self):
        """Polls NVML every ~10 ms: the timed region of the default run is ~120 ms, which `nvidia-smi -lms 100` sees
        once.  Any failure simply ends this thread; the nvidia-smi rows remain."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(self.gpu))
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.nvml_stop.is_set():
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:                                       # noqa: BLE001
                    power = 0.0
                self.nvml_rows.append((sm, smax, power, int(get_reasons(h))))
                time.sleep(0.01)
        except Exception:                                               # noqa: BLE001
            return

    def start(self):
        self.nvml_thread = threading.Thread(target=self._nvml_pump, daemon=True)
        self.nvml_thread.start()
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", self.gpu], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    # NVML clocks-event-reason bits (nvml.h): sw_power_cap 0x4, hw_slowdown 0x8, sw_thermal 0x20, hw_thermal 0x40
    NVML_BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
                 ("hw_thermal_slowdown", 0x40))

    def stop(self):
        self.nvml_stop.set()
        if self.nvml_thread is not None:
            self.nvml_thread.join(timeout=2)
        smi = self._stop_smi()
        rows = list(self.nvml_rows)
        if len(rows) < 2:
            return smi
        reasons = set(r for r in smi.get("reasons", []) if r in dict(self.NVML_BITS))
        for _, _, _, bits in rows:
            reasons.update(name for name, bit in self.NVML_BITS if bits & bit)
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": max(r[1] for r in rows),
                "power_w_max": max(r[2] for r in rows), "samples": len(rows), "source": "nvml, 10 ms period",
                "reasons": sorted(reasons)}

    def _stop_smi(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(scale=20, mmd_times_full=5, threads=None):
    """One bounded sample of the reference arm: an oracle A2GNN train step
    (forward_model + zero_grad + backward + Adam.step, pygda/models/a2gnn.py:314-319) on a
    1/scale-size pair of graphs with ONE MMD sample, plus a stand-alone timing of that MMD
    sample; full-step estimate = scale * (t_step - t_mmd) + 5 * t_mmd  (graph terms scale
    linearly in N and nnz; the MMD term is size-independent: n = 2000 rows, 5 samples)."""
    import torch
    from oracle import mmd as OM
    from oracle.data import Data as OData
    from oracle.models import A2GNN as OracleA2GNN
    from pygda_b200.synthetic import domain_pair

    if threads:
        torch.set_num_threads(threads)
    n, e = CFG["nodes"] // scale, CFG["edges"] // scale
    st = cpu_reference_sample.__dict__.setdefault("state", {})
    if not st:
        src, tgt = domain_pair(n, e, CFG["feats"], CFG["classes"], seed=0)
        torch.manual_seed(0)
        est = OracleA2GNN(CFG["feats"], CFG["hid"], CFG["classes"], num_layers=CFG["layers"],
                          dropout=CFG["dropout"], s_pnums=CFG["s_pnums"], t_pnums=CFG["t_pnums"],
                          weight=CFG["weight"], weight_decay=CFG["weight_decay"], lr=CFG["lr"],
                          epoch=CFG["epochs"], device="cpu")
        st.update(src=OData(x=src.x, edge_index=src.edge_index, y=src.y),
                  tgt=OData(x=tgt.x, edge_index=tgt.edge_index, y=tgt.y), est=est)
    est = st["est"]
    est.mmd_indices = OM.draw_mmd_indices(n, n, 1000, 1)
    t0 = time.perf_counter()
    est.train_step(st["src"], st["tgt"], epoch=0)
    t_step = time.perf_counter() - t0
    a = torch.randn(1000, CFG["hid"], requires_grad=True)
    b = torch.randn(1000, CFG["hid"], requires_grad=True)
    t0 = time.perf_counter()
    OM.get_mmd(a, b).backward()              # reference-faithful n x n x d broadcast (mmd.py:44-46)
    t_mmd = time.perf_counter() - t0
    full = scale * max(t_step - t_mmd, 0.0) + mmd_times_full * t_mmd
    # The same MMD sample with the [n, n, d] temporary replaced by the Gram form |a|^2 + |b|^2 - 2ab (same values to
    # fp32 round-off): what the CPU arm would cost if the reference's MMD were written sanely -- reported NEXT to the
    # reference-faithful number so that the speed-up is not inflated by that temporary.
    try:
        def gram(total):
            sq = (total * total).sum(1)
            return (sq[:, None] + sq[None, :] - 2.0 * total @ total.t()).clamp_(min=0)
        a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        t0 = time.perf_counter()
        OM.get_mmd(a2, b2, sqdist=gram).backward()
        t_gram = time.perf_counter() - t0
        st["full_with_gram_mmd"] = scale * max(t_step - t_mmd, 0.0) + mmd_times_full * t_gram
        st["t_mmd_gram"] = t_gram
    except Exception:                                    # noqa: BLE001 -- an extra, never fatal
        st.pop("full_with_gram_mmd", None)
    return full, t_step, t_mmd


def _gram_note():
    """cpu_baseline extras: the same estimate with a memory-sane (Gram-form) MMD on the CPU."""
    st = cpu_reference_sample.__dict__.get("state", {})
    if "full_with_gram_mmd" not in st:
        return {}
    return {"value_with_gram_mmd_estimate": 1.0 / st["full_with_gram_mmd"],
            "gram_mmd_note": "same sample with the reference's [n,n,d] MMD temporary replaced by the Gram form on the "
                             "CPU (t_mmd=%.3fs): the reference-faithful `value` is dominated by that temporary" % st["t_mmd_gram"]}


def calibrate_threads(scale, cores):
    """torch's CPU scatter_add / index_select path does not scale to very wide hosts (128 threads
    were 8x slower than 8 on the B200 box); give the reference its best thread count."""
    import torch
    best, best_t = cores, None
    for n in sorted({min(cores, c) for c in (8, 16, 32, cores)}):
        torch.set_num_threads(n)
        cpu_reference_sample(scale)                  # warm-up at this setting
        full, _, _ = cpu_reference_sample(scale)
        if best_t is None or full < best_t:
            best, best_t = n, full
    torch.set_num_threads(best)
    return best


class FullScaleReference:
    """The reference arm proper: the oracle's A2GNN train step (forward_model + zero_grad + backward + Adam.step,
    pygda/models/a2gnn.py:314-319) on the FULL config-2 graph pair (100k nodes / 1M edges / 6775 features per
    domain, five reference-faithful [n, n, d] MMD samples per step) on the host cores -- measured, not extrapolated."""

    def __init__(self, src=None, tgt=None, scale=1):
        import torch
        from oracle.data import Data as OData
        from oracle.models import A2GNN as OracleA2GNN
        from pygda_b200.synthetic import domain_pair
        self.nodes, self.edges = CFG["nodes"] // scale, CFG["edges"] // scale
        if src is None:
            src, tgt = domain_pair(self.nodes, self.edges, CFG["feats"], CFG["classes"], seed=0)
        torch.manual_seed(0)
        self.est = OracleA2GNN(CFG["feats"], CFG["hid"], CFG["classes"], num_layers=CFG["layers"],
                               dropout=CFG["dropout"], s_pnums=CFG["s_pnums"], t_pnums=CFG["t_pnums"],
                               weight=CFG["weight"], weight_decay=CFG["weight_decay"], lr=CFG["lr"],
                               epoch=CFG["epochs"], device="cpu")
        self.src = OData(x=src.x, edge_index=src.edge_index, y=src.y)
        self.tgt = OData(x=tgt.x, edge_index=tgt.edge_index, y=tgt.y)
        self.n = 0

    def step(self):
        t0 = time.perf_counter()
        self.est.train_step(self.src, self.tgt, epoch=self.n % CFG["epochs"])
        self.n += 1
        return time.perf_counter() - t0


REFERENCE_BUDGET_S = 420.0      # warm-up + timed steps of the reference arm (the graphs take ~20 s more to generate)


def run_reference_arm(args, rank, world):
    """`--impl reference`: FULL-SCALE oracle steps on this box's host cores.  Runs the requested warm-up and step
    counts when they fit REFERENCE_BUDGET_S, else as many as fit (never fewer than 1 warm-up + 2 timed steps);
    `steps` / `warmup` in the line are what was actually run, `ms_per_step` the measured mean."""
    if rank != 0:
        return
    cores = calibrate_threads(20, os.cpu_count() or 1)          # thread count at which the oracle is fastest
    small_full, small_step, small_mmd = cpu_reference_sample(20)
    ref = FullScaleReference(scale=args.ref_scale)
    t_first = ref.step()                                         # warm-up 1 (allocator, page faults)
    want_w, want_k = max(args.warmup, 1), max(args.steps, 1)
    warm, steps = want_w, want_k
    if (want_w - 1 + want_k) * t_first > REFERENCE_BUDGET_S:
        warm = 1
        steps = int(min(want_k, max(2, (REFERENCE_BUDGET_S - t_first) // max(t_first, 1e-9))))
    for _ in range(warm - 1):
        ref.step()
    times = [ref.step() for _ in range(steps)]
    full = statistics.mean(times)
    value = 1.0 / full
    sample = ("%s: oracle A2GNN train step on the whole config-2 graph pair (%d nodes, %d edges, F=%d per "
              "domain; 5 reference-faithful [2000,2000,%d] MMD samples per step); %d warm-up + %d timed steps "
              "(requested %d + %d; budget %.0f s), mean %.2f s, min %.2f s, max %.2f s"
              % ("full scale" if args.ref_scale == 1 else "REDUCED 1/%d scale (test knob --ref-scale)" % args.ref_scale,
                 ref.nodes, ref.edges, CFG["feats"], CFG["hid"], warm, steps, want_w, want_k,
                 REFERENCE_BUDGET_S, full, min(times), max(times)))
    cfg = {"workload": workload_name(), "device": "cpu"}
    if world > 1:
        # The N-GPU arm counts config-2-sized graph-epochs per second (one community per GPU).  A host works through
        # N such graphs one after the other, so its rate in the same unit is the single-graph rate measured here.
        cfg["parallelism"] = ("per-graph rate: the %d-GPU arm's value counts config-2-sized graph-epochs per second; "
                              "the host processes such graphs sequentially at the rate measured on one" % world)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": full * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                  "estimated": False,
                                  "small_scale_estimate": {
                                      "value": 1.0 / small_full, "estimated": True,
                                      "how": "round-1 estimate, kept as a side key only: 1/20-scale graph pair with one MMD "
                                             "sample, 20*(t_step - t_mmd) + 5*t_mmd; t_step=%.2fs t_mmd=%.2fs"
                                             % (small_step, small_mmd)}},
                                 **_gram_note()),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def data_bytes(d):
    import torch
    return sum(v.numel() * v.element_size() for v in d.__dict__.values() if torch.is_tensor(v))


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pygda_b200 import ops
    from pygda_b200._lib import load
    from pygda_b200.graph import graph_for
    from pygda_b200.models import A2GNN
    from pygda_b200.optim import Adam
    from pygda_b200.data import Data
    from pygda_b200.synthetic import domain_pair

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU arm")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    distributed = world > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    lib = load()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    hp = dict(in_dim=CFG["feats"], hid_dim=CFG["hid"], num_classes=CFG["classes"], mode="node",
              num_layers=CFG["layers"], dropout=CFG["dropout"], s_pnums=CFG["s_pnums"], t_pnums=CFG["t_pnums"],
              adv=False, weight=CFG["weight"], weight_decay=CFG["weight_decay"], lr=CFG["lr"],
              epoch=CFG["epochs"], device=str(dev), verbose=0)
    torch.manual_seed(0)
    run_epoch = None
    if not distributed:
        src, tgt = domain_pair(CFG["nodes"], CFG["edges"], CFG["feats"], CFG["classes"], seed=0, device=dev)
        model = A2GNN(**hp)
        if args.no_cuda_graph:
            model.cuda_graph = False
        # exactly what fit() does before its epoch loop (pygda_b200/models/a2gnn.py: prepare_fit): loaders, model,
        # optimiser and -- the default for full-batch node mode -- the CUDA-graph step; epoch 0 ran eagerly inside
        run_epoch = model.prepare_fit(src, tgt)
        s_batch, t_batch = next(iter(model.source_loader)), next(iter(model.target_loader))
        parallelism = "single GPU"
    else:
        # Weak scaling over a 1-D node partition: the global source / target graphs
        # are `world` citation-shaped communities of the config-2 size with a 5 % edge cut; rank r owns
        # community r (rows, features, labels).  Neighbour rows owned by other GPUs are gathered over
        # NVLink inside the aggregation kernel; weight gradients are all-reduced once per step.
        from pygda_b200.dist import PeerGroup, attach_partition
        from pygda_b200.models.dist_a2gnn import DistA2GNN
        from pygda_b200.synthetic import community_block
        group = PeerGroup(device=dev)
        src = community_block(world, rank, CFG["nodes"], CFG["edges"], CFG["feats"], CFG["classes"], seed=0,
                              device=dev)
        tgt = community_block(world, rank, CFG["nodes"], CFG["edges"], CFG["feats"], CFG["classes"], seed=1,
                              feature_shift=1.4, degree_offset=48.0, device=dev)
        src, tgt = attach_partition(src, group), attach_partition(tgt, group)
        model = DistA2GNN(group=group, **hp)
        if args.no_cuda_graph:
            model.cuda_graph = False
        graph_error = None
        try:
            # the same prepare_fit as on one GPU: the partitioned step (peer aggregations with their device-side
            # barriers, NCCL all-reduces) captured as ONE CUDA graph per rank and replayed
            run_epoch = model.prepare_fit(src, tgt)
            s_batch, t_batch = next(iter(model.source_loader)), next(iter(model.target_loader))
        except Exception as exc:                      # noqa: BLE001 -- measure the eager path and say why
            graph_error = repr(exc)[:300]
            ok = torch.tensor([0], device=dev)
        else:
            ok = torch.tensor([1 if (run_epoch is not None or args.no_cuda_graph) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # every rank takes the same path
        if int(ok.item()) == 0 or run_epoch is None:
            run_epoch = None
            model.cuda_graph = False
            if not hasattr(model, "optimizer") or graph_error is not None:
                model.prepare_fit(src, tgt)
            s_batch, t_batch = next(iter(model.source_loader)), next(iter(model.target_loader))
        parallelism = ("1-D node partition: %d communities of %dk nodes (one per GPU, 5%% cross-partition edges), "
                       "NVLink peer gathers in the aggregation kernel, NCCL gradient all-reduce; value counts "
                       "config-2-sized graph-epochs per second" % (world, CFG["nodes"] // 1000))
    sopt = model.optimizer
    step_no = [1]

    def one_step(sb, tb):
        alpha = model.alpha_at(step_no[0] % CFG["epochs"], CFG["epochs"])
        step_no[0] += 1
        loss, _, _, _ = model.train_step(sb, tb, alpha, sopt)
        return loss

    # ---------------- device-resident throughput (`value`) ----------------
    # Single GPU: fit()'s default path -- the loop body replayed from a CUDA graph (pygda_b200/models/graphed.py),
    # the same kernels on the same resident buffers, MMD indices still drawn on the CPU generator and staged in
    # before every replay.  `--no-cuda-graph` / N > 1: every kernel issued eagerly.
    def timed_step(sb, tb):
        if run_epoch is not None:
            e = step_no[0]
            step_no[0] += 1
            return run_epoch(e % CFG["epochs"] or 1)[0]
        return one_step(sb, tb)

    warm = args.warmup if args.skip_e2e else max(args.warmup, 3)      # profiling runs may use fewer
    for _ in range(warm):
        timed_step(s_batch, t_batch)
    barrier()
    gstep = getattr(model, "graphed_step", None) if run_epoch is not None else None
    graph_note = ("CUDA graph replay (%d kernels per step), fit()'s default" % gstep.launches_per_replay
                  if gstep is not None else "eager launches")
    if distributed and gstep is None and locals().get("graph_error"):
        graph_note += " (graph capture failed: %s)" % graph_error
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.gda_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures the timed steps only
    ev0.record()
    host_t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = timed_step(s_batch, t_batch)
    host_issue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps     # CPU time to ISSUE a step
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = lib.gda_launch_count() - launches0
    if gstep is not None:                 # replayed kernels do not pass through the library's launch counter
        launches += gstep.launches_per_replay * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps / (ms_total / 1e3)
    final_loss = float(loss.item())

    # ---------------- roofline of the aggregation kernel (instrumented repeat) ----------------
    ops.PROFILE = []
    overlap, model.overlap_streams = model.overlap_streams, False      # time the kernel alone, not co-scheduled
    for _ in range(min(args.steps, 5)):
        one_step(s_batch, t_batch)
    torch.cuda.synchronize()
    model.overlap_streams = overlap
    recs, ops.PROFILE = ops.PROFILE, None
    if distributed:
        tg = t_batch.edge_index._gda_partition.graph(t_batch.edge_index, 1 | 4)      # SELF_LOOPS | NORM_SYM_COL
        nnz = tg.local_nnz
    else:
        tg = graph_for(t_batch.edge_index, CFG["nodes"])
        nnz = tg.nnz
    n_nodes, H = CFG["nodes"], CFG["hid"]
    b_alg = 4 * (n_nodes + 1) + 8 * nnz + 2 * 4 * n_nodes * H
    def launches_of(nb_):
        by_kind = {}
        for (a, b, meta) in recs:
            if meta[:4] == (n_nodes, H, "float32", nb_):
                by_kind.setdefault(meta[4], []).append(a.elapsed_time(b))
        if not by_kind:
            return None, []
        kind = max(by_kind, key=lambda k_: len(by_kind[k_]))        # the kernel that carries the step
        return kind, by_kind[kind]

    n_prof = max(min(args.steps, 5), 1)
    kind1, times = launches_of(1)
    spmm_ms = statistics.mean(times) if times else float("nan")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "spmm_traffic.json")))
    except Exception:
        pass
    # The factored (unit-weight) kernel reads no edge weights: 4 bytes per non-zero fewer, + the dinv vector.
    def alg_bytes(kind, nb_):
        idx = 4 * (n_nodes + 1) + (4 * nnz + 4 * n_nodes if str(kind).startswith("unit-weight") else 8 * nnz)
        return idx + nb_ * 2 * 4 * n_nodes * H
    kname = {"unit-weight": "k_spmm_unw<4,16,...> (factored D S D form: no per-edge weights)",
             "unit-weight-halo": "k_spmm_unw<4,...> on the rank's rectangular block (factored D S D form; halo rows "
                                 "pushed by their owners after every step)",
             "weighted": "k_spmm_tasks<float,4,4,...>"}
    b_alg = alg_bytes(kind1, 1)
    achieved = b_alg / (spmm_ms * 1e-3) / 1e9
    per_step_spmm = len(times) / n_prof
    # layer >= 2 of the two bottleneck evaluations per domain runs as ONE launch over the stacked pair (nb = 2):
    # same index bytes, twice the feature bytes
    kind2, times2 = launches_of(2)
    batched = None
    if times2:
        ms2 = statistics.mean(times2)
        b_alg2 = alg_bytes(kind2, 2)
        batched = {"kernel": "%s, NB=2 (two stacked [N,128] matrices per launch)" % kname.get(kind2, kind2),
                   "alg_bytes_per_launch": b_alg2, "us_per_launch": ms2 * 1e3,
                   "achieved": b_alg2 / (ms2 * 1e-3) / 1e9, "frac": b_alg2 / (ms2 * 1e-3) / 1e9 / peak,
                   "launches_per_step": len(times2) / n_prof}
    gather_bytes = 4 * (n_nodes + 1) + (4 if str(kind1).startswith("unit-weight") else 8) * nnz + 4 * nnz * H + 4 * n_nodes * H
    roofline = {"bound": "hbm", "kernel": "%s (A_hat x, H=128, N=100k, nnz=%d)" % (kname.get(kind1, kind1), nnz),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "traffic": ((traffic or {}).get(kind1) or {}).get("bytes_per_launch"),
                "traffic_source": ((traffic or {}).get(kind1) or {}).get("source"),
                "gather_bytes_per_launch": gather_bytes, "gathered_GBps": gather_bytes / (spmm_ms * 1e-3) / 1e9,
                "l2_gather_ceiling_GBps": 15000.0,
                "l2_gather_ceiling_source": "profiles/probes/gather_probe6 (regular 12-gather rows, 64 warps/SM: 14.9-15.3 "
                                            "TB/s of gathered rows on this GPU; the feature matrix is L2-resident, so the "
                                            "kernel is bounded by the L2 gather path, not by HBM)",
                "alg_bytes_per_launch": b_alg, "us_per_launch": spmm_ms * 1e3,
                "launches_per_step": per_step_spmm,
                "share_of_step": (per_step_spmm * spmm_ms + (batched["launches_per_step"] * batched["us_per_launch"] * 1e-3
                                                              if batched else 0.0)) / (ms_total / args.steps),
                "batched": batched,
                "how": "CUDA events around each aggregation launch in an instrumented repeat of the timed steps"}

    # layer-1 products from the tile-packed sparse X (csrc/gemm_xt.cu): tensor-bound, not HBM-bound -- the record says
    # how far from the dense-bf16 tensor peak the three-term split runs and how few bytes it moves
    def xt_record(kind):
        ts = [(a.elapsed_time(b), meta) for (a, b, meta) in recs if len(meta) == 5 and meta[2] == kind]
        if not ts:
            return None
        ms_ = statistics.mean(t_ for t_, _ in ts)
        rows_, cols_, _, n_, nnz_ = ts[0][1]
        flops = 3 * 2.0 * rows_ * cols_ * n_                       # three bf16 UMMA terms per product (fp32-accurate)
        bytes_ = 5 * nnz_ + 20 * (-(-rows_ // 32)) * (-(-cols_ // 64)) + 4 * rows_ * n_ * (2 if kind == "xt_dw" else 1) \
            + 4 * cols_ * n_
        tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        return {"kernel": "k_gemm_xt<%s> (X [%d x %d], %.1f %% non-zero, tile-packed; N = %d)"
                          % ("dW = G^T X" if kind == "xt_dw" else "X W^T", rows_, cols_, 100.0 * nnz_ / (rows_ * cols_), n_),
                "bound": "tensor", "us_per_launch": ms_ * 1e3, "launches_per_step": len(ts) / n_prof,
                "tensor_flops_per_launch": flops, "achieved": flops / (ms_ * 1e-3) / 1e12, "peak": tpeak,
                "unit": "TFLOP/s", "frac": flops / (ms_ * 1e-3) / 1e12 / tpeak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 8192^3, back to back)",
                "alg_bytes_per_launch": bytes_, "dense_operand_bytes_it_replaces": 4 * rows_ * cols_,
                "note": "useful fp32 flops are a third of the tensor flops (Ah*Bh + Ah*Bl + Al*Bh); the dense "
                        "split-bf16 kernel it replaces streams 4 bytes per ELEMENT of X and is HBM-bound"}
    gemm_rec = {k_: v_ for k_, v_ in (("forward", xt_record("xt_fwd")), ("weight_gradient", xt_record("xt_dw"))) if v_}
    if gemm_rec:
        roofline["gemm"] = gemm_rec

    # ---------------- end to end from pinned host buffers (`e2e`) ----------------
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_total / args.steps,
                              "host_issue_ms_per_step": host_issue_ms, "gpu_launches": int(launches),
                              "issue": graph_note, "roofline": roofline, "note": "profiling run"}), flush=True)
        return
    # Host copies of the step's inputs; every timed step copies x / edge_index / y of both graphs host->device and
    # reads the loss back.  `Data.pin_memory()` (what fit()'s loaders do) keeps a sparse fp32 x ROW-COMPRESSED in pinned
    # memory: only the non-zeros cross PCIe, the dense matrix is rebuilt on the GPU bit for bit (gda_unpack_rows_f32).
    #   e2e.value          fit()'s DEFAULT path (A2GNN.prepare_fit on the host graphs): CUDA-graph replay, the copy of
    #                      epoch e+1 double-buffered behind the replay of epoch e (models/graphed.py: StagedBatch)
    #   e2e.eager_serial   the reference's schedule: copy, then compute, every kernel issued from Python
    #                      (cuda_graph = False, prefetch = False) -- round 1's e2e.value
    #   e2e.dense_staging  eager_serial with the plain dense pinned copy (pack=False)
    src_h, tgt_h = src.to("cpu"), tgt.to("cpu")
    del src, tgt, s_batch, t_batch, gstep, run_epoch
    if hasattr(model, "graphed_step"):
        del model.graphed_step
    torch.cuda.empty_cache()
    e2e_steps = max(3, min(args.steps, 20))

    def timed_loop(step_fn, n):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n):
            step_fn(i).item()                          # loss read-back every step
        t1.record()
        barrier()
        t = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n / (float(t.item()) / 1e3)

    e2e_note, e2e_eager, e2e_dense, h2d_dense = "eager launches, copy then compute", None, None, None
    staged_ok, e2e_fallback = True, None
    if distributed:
        # the same default fit() path over the partition: every rank stages its own row block (host -> its GPU over its own
        # PCIe link) behind the replay of the captured partitioned step; all ranks must agree on the path
        model.cuda_graph = not args.no_cuda_graph
        try:
            run_h = model.prepare_fit(src_h, tgt_h)
            flag = torch.tensor([1 if run_h is not None else 0], device=dev)
        except Exception as exc:                      # noqa: BLE001 -- fall back to the eager schedule and say why
            run_h, flag = None, torch.tensor([0], device=dev)
            e2e_fallback = repr(exc)[:200]
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        staged_ok = int(flag.item()) == 1
        if not staged_ok:
            run_h = None
            model.cuda_graph = False
            if hasattr(model, "graphed_step"):
                del model.graphed_step
            torch.cuda.empty_cache()
    if not distributed or staged_ok:
        if not distributed:
            model.cuda_graph = not args.no_cuda_graph
            run_h = model.prepare_fit(src_h, tgt_h)        # pins (tile-packs) the host graphs like fit() does
        sopt = model.optimizer
        sb_p, tb_p = next(iter(model.source_loader)), next(iter(model.target_loader))
        h2d = sb_p.h2d_nbytes() + tb_p.h2d_nbytes()
        packed = "_packed_x" in sb_p.__dict__
        if run_h is not None:
            gs = model.graphed_step
            h2d = gs.h2d_bytes_per_step                 # counted from the tensors StagedBatch copies every epoch
            tiled = all(sb is not None and sb.tiles is not None for sb in gs.staged)
            e2e_note = ("fit() default: CUDA-graph replay (%d kernels), next epoch's host->device copy "
                        "double-buffered behind it; %s" % (
                            gs.launches_per_replay,
                            "x crosses PCIe TILE-PACKED with exponent-packed values (4.65 bytes per non-zero: 3 bytes of "
                            "sign + mantissa, a 4-bit exponent code, a position byte; lossless) into the buffers the first "
                            "layer's tensor-core GEMMs read: no dense rebuild, no operand split" if tiled else
                            "consumed by unpack + operand-split kernels before each replay"))
            ep = [1]

            def graphed_h(i):
                ep[0] += 1
                return run_h(ep[0] % CFG["epochs"] or 1)[0]
            for i in range(3):
                graphed_h(i).item()
            e2e_value = timed_loop(graphed_h, e2e_steps)
            del run_h, gs
            del model.graphed_step
            torch.cuda.empty_cache()
        else:
            for i in range(3):
                one_step(sb_p, tb_p).item()
            e2e_value = timed_loop(lambda i: one_step(sb_p, tb_p), e2e_steps)
        if distributed:
            del sb_p, tb_p
        # the reference's serial schedule on the same pinned batches (side keys, one GPU only)
        n_side = max(3, min(args.steps, 10))
        for i in range(3 if not distributed else 0):
            one_step(sb_p, tb_p).item()
        if not distributed:
          e2e_eager = timed_loop(lambda i: one_step(sb_p, tb_p), n_side)
          sb_d = Data(x=sb_p.x, edge_index=sb_p.edge_index, y=sb_p.y).pin_memory(pack=False)
          tb_d = Data(x=tb_p.x, edge_index=tb_p.edge_index, y=tb_p.y).pin_memory(pack=False)
          for a_, b_ in ((sb_d, sb_p), (tb_d, tb_p)):          # same graph-cache identity as the packed form
              for attr in ("_gda_partition", "_gda_key", "_gda_keepalive"):
                  if hasattr(b_.edge_index, attr):
                      setattr(a_.edge_index, attr, getattr(b_.edge_index, attr))
          del sb_p, tb_p
          h2d_dense = sb_d.h2d_nbytes() + tb_d.h2d_nbytes()
          for i in range(2):
              one_step(sb_d, tb_d).item()
          e2e_dense = timed_loop(lambda i: one_step(sb_d, tb_d), max(3, n_side // 2))
          del sb_d, tb_d
    else:
        sb_p = src_h if "_packed_x" in src_h.__dict__ or src_h.x.is_pinned() else src_h.pin_memory()
        tb_p = tgt_h if "_packed_x" in tgt_h.__dict__ or tgt_h.x.is_pinned() else tgt_h.pin_memory()
        h2d = sb_p.h2d_nbytes() + tb_p.h2d_nbytes()
        packed = "_packed_x" in sb_p.__dict__
        for i in range(3):
            one_step(sb_p, tb_p).item()
        e2e_value = timed_loop(lambda i: one_step(sb_p, tb_p), e2e_steps)
        del sb_p, tb_p
        if e2e_fallback:
            e2e_note += " (the staged graph-replay path was not taken: %s)" % e2e_fallback

    # ---------------- the other BASELINE configurations, as side measurements (`other_configs`) ----------------
    # config 3 (UDAGCN 1M / 10M, bf16, one GPU) rides along at N = 1, config 4 (GRADE-MMD 5M / 50M, random partition) at
    # N = 4, config 5 (AdaGCN graph-level, 50k graphs / batch 512) at N = 8 -- the GPU counts BASELINE.json names them on.
    # `python bench.py --config K --gpus N` runs any of them alone at any N.  Never allowed to take the main line down.
    other = {}
    if not args.no_other_configs:
        try:
            del model, sopt
            from pygda_b200.graph import clear_graph_cache
            clear_graph_cache()
            ops.split_cache.clear(); ops.bf16_cache.clear()
            torch.cuda.empty_cache()
            if world == 1:
                other["config3"] = run_config3(dev, steps=5, warmup=2)
            elif world == 4:
                other["config4"] = run_config4(dev, rank, world, group, steps=3, warmup=1)
            elif world == 8:
                other["config5"] = run_config5(dev, rank, world, group, steps=20, warmup=3)
        except Exception as exc:                                            # noqa: BLE001
            other["error"] = repr(exc)[:400]
    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "hid": CFG["hid"], "layers": CFG["layers"],
                       "s_pnums": CFG["s_pnums"], "t_pnums": CFG["t_pnums"], "dropout": CFG["dropout"],
                       "mmd_weight": CFG["weight"], "optimizer": "Adam lr=0.01 wd=0.005",
                       "epoch_definition": "full-batch: 1 epoch = 1 optimiser step; F1/logging excluded",
                       "parallelism": parallelism, "issue": graph_note,
                       "l2_policy": "working set larger than L2 (tile-packed x, 0.24 GB per domain, streamed 4x per "
                                    "step; ~20 fp32 [100k,128] activation matrices of 51 MB written and re-read); "
                                    "no explicit flush",
                       "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                    "h2d_bytes_per_step_per_gpu": h2d,
                    "steps": e2e_steps, "how": e2e_note,
                    "staging": ("pinned host inputs, x tile-packed (lossless; only its non-zeros cross PCIe: 4.65 bytes per "
                                "non-zero + 20 bytes per 32 x 64 sub-tile)" if packed else "pinned host inputs, dense"),
                    "eager_serial": ({"value": e2e_eager, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                      "how": "cuda_graph=False, prefetch=False: copy, then compute (round 1's e2e.value)"}
                                     if e2e_eager is not None else None),
                    "dense_staging": ({"value": e2e_dense, "unit": UNIT, "h2d_bytes_per_step": h2d_dense}
                                      if e2e_dense is not None else None)},
            "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms, "roofline": roofline,
            "clocks": clocks, "other_configs": other or None}
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of the SAME workload: full-scale oracle steps on the graphs the GPU arm just trained on
        # (one warm-up + one timed step; the warm-up itself when a step takes longer than 20 s)
        cores = calibrate_threads(20, os.cpu_count() or 1)
        ref = FullScaleReference(src_h, tgt_h)
        t_first = ref.step()
        t_full, what = (ref.step(), "1 warm-up + 1 timed step") if t_first < 20.0 else (t_first, "one cold step")
        line["cpu_baseline"] = {
            "value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "port", "estimated": False,
            "sample": "full scale: oracle train step on the whole config-2 graph pair (%d nodes, %d edges, F=%d per "
                      "domain, 5 reference-faithful MMD samples); %s, %.2f s (first step %.2f s)"
                      % (CFG["nodes"], CFG["edges"], CFG["feats"], what, t_full, t_first)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ BASELINE configs 3 / 4 / 5 (parity-test cases)
def _peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


def _timed_steps(step, warmup, steps, dev, distributed):
    """`warmup` untimed + `steps` timed calls of step(i); ms per step, max over ranks (CUDA events)."""
    import torch
    import torch.distributed as dist
    for i in range(warmup):
        step(i)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = step(warmup + i)
    e1.record()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps, out


def _aggregation_profile(step, n_steps, match):
    """Per-launch CUDA-event times of the aggregation launches whose profile record satisfies `match(meta)`."""
    import torch
    from pygda_b200 import ops
    ops.PROFILE = []
    for i in range(n_steps):
        step(10_000 + i)
    torch.cuda.synchronize()
    recs, ops.PROFILE = ops.PROFILE, None
    return [a.elapsed_time(b) for a, b, meta in recs if match(meta)]


def run_config3(dev, steps=10, warmup=3, nodes=1_000_000, edges=10_000_000):
    """BASELINE.json configs[2]: UDAGCN (GRL + discriminator, ppmi=False), per domain 1M nodes / 10M directed edges /
    512 features / 5 classes, hid 256, 2 layers, bf16 feature path (fp32 parameters / accumulation / losses), one B200;
    hyper-parameters of benchmark/node/run_citation.sh:90 (lr 1e-4, wd 1e-3, 400 epochs)."""
    import itertools
    import torch
    from pygda_b200.data import Data
    from pygda_b200.graph import Graph, clear_graph_cache
    from pygda_b200.models import UDAGCN
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import bow_features, powerlaw_edge_index_device
    F_, H, C = 512, 256, 5
    g = torch.Generator().manual_seed(3)

    def domain(seed, shift, offset):
        x = bow_features(nodes, F_, seed=seed + 1000, shift=shift, device=dev)
        x._gda_const = True
        return Data(x=x, edge_index=powerlaw_edge_index_device(nodes, edges, seed=seed, offset=offset, device=dev),
                    y=torch.randint(C, (nodes,), generator=g).to(dev))
    src, tgt = domain(30, 1.5, 32.0), domain(31, 1.4, 48.0)
    torch.manual_seed(0)
    est = UDAGCN(in_dim=F_, hid_dim=H, num_classes=C, num_layers=2, ppmi=False, lr=1e-4, weight_decay=1e-3,
                 epoch=400, device=str(dev), verbose=0, feature_dtype=torch.bfloat16)
    est.udagcn = est.init_model()
    opt = Adam(itertools.chain(*[m.parameters() for m in est.udagcn.models]), lr=1e-4, weight_decay=1e-3)
    step = lambda i: est.train_step(src, tgt, min((i % 400 + 1) / 400.0, 0.05), i % 400, opt)[0]   # noqa: E731
    ms, loss = _timed_steps(step, warmup, steps, dev, False)
    tm = _aggregation_profile(step, 3, lambda m: m[:3] == (nodes, H, "bfloat16"))
    gr = Graph(tgt.edge_index, nodes, None, 1 | 8)               # SELF_LOOPS | NORM_SYM_ROW, as CachedGCNConv builds it
    b_alg = 4 * (nodes + 1) + 8 * gr.nnz + 2 * 2 * nodes * H
    us = statistics.mean(tm) * 1e3 if tm else float("nan")
    peak, src_ = _peak_hbm()
    out = {"config": 3, "workload": "UDAGCN synthetic %dM nodes / %dM edges / %d feat, hid %d, bf16 features, 1 x B200 "
                                    "(BASELINE.json configs[2])" % (nodes // 10**6, edges // 10**6, F_, H),
           "metric": "udagcn_train_epochs_per_sec", "value": 1e3 / ms, "unit": "epochs/s", "ms_per_step": ms,
           "steps": steps, "warmup": warmup, "dtype": "bf16 features, f32 parameters/accumulation", "loss": float(loss),
           "parity_note": "bf16 feature path: outputs within 1e-2, gradients by Frobenius norm (tests/test_gpu_bf16.py); "
                          "the 1e-4 bar applies to the fp32 path",
           "roofline": {"bound": "hbm", "kernel": "k_spmm_tasks<bf16,8,...> (H=256 bf16, N=%d, nnz=%d)" % (nodes, gr.nnz),
                        "alg_bytes_per_launch": b_alg, "us_per_launch": us, "achieved": b_alg / us / 1e3, "peak": peak,
                        "unit": "GB/s", "frac": b_alg / us / 1e3 / peak, "peak_source": src_,
                        "gathered_GBps": 2 * gr.nnz * H / us / 1e3, "launches_per_step": len(tm) / 3.0,
                        "working_set_note": "1.1 GB per launch: far beyond the 126 MB L2, the gathers are served by HBM"},
           "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    del est, opt, src, tgt, gr
    clear_graph_cache()
    from pygda_b200 import ops
    ops.split_cache.clear(); ops.bf16_cache.clear()
    torch.cuda.empty_cache()
    return out


def run_config4(dev, rank, world, group, steps=5, warmup=2, nodes=5_000_000, edges=50_000_000):
    """BASELINE.json configs[3]: GRADE (Gaussian-MMD), per domain 5M nodes / 50M directed edges / 256 features / 5
    classes, hid 128, 2 layers (grade.py:68 default; run_citation.sh:85 uses 5), fp32, 1-D node partition over `world`
    GPUs of a RANDOM graph (node ids shuffled: no locality, (world-1)/world of the gathered rows are remote)."""
    import torch
    from pygda_b200.data import Data
    from pygda_b200.graph import clear_graph_cache
    from pygda_b200.models import GRADE
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import bow_features, powerlaw_edge_index_device
    F_, H, C = 256, 128, 5
    hp = dict(in_dim=F_, hid_dim=H, num_classes=C, num_layers=2, dropout=0.5, disc="MMD", weight=0.002,
              weight_decay=0.0, lr=0.001, epoch=300, device=str(dev), verbose=0)
    distributed = world > 1
    lo, hi = (0, nodes)
    if distributed:
        from pygda_b200.dist import attach_partition, block_range
        from pygda_b200.models.dist_grade import DistGRADE
        lo, hi = block_range(nodes, world, rank)
    gl = torch.Generator().manual_seed(4)

    def domain(seed, shift, offset):
        ei = powerlaw_edge_index_device(nodes, edges, seed=seed, offset=offset, device=dev)
        x = bow_features(hi - lo, F_, seed=seed + 1000 + rank, shift=shift, device=dev)
        x._gda_const = True
        y = torch.randint(C, (nodes,), generator=gl)[lo:hi].to(dev)
        d = Data(x=x, edge_index=ei, y=y)
        if distributed:
            d.num_nodes_global, d.row_lo, d.row_hi = nodes, lo, hi
            d = attach_partition(d, group)
        return d
    src, tgt = domain(40, 1.5, 32.0), domain(41, 1.4, 48.0)
    rpr = (nodes + world - 1) // world
    remote_frac = float(((tgt.edge_index[0] // rpr) != (tgt.edge_index[1] // rpr)).float().mean()) if distributed else 0.0
    # non-zeros of this rank's row block whose column lives on another GPU (target graph; self loops are local)
    own = (tgt.edge_index[1] >= lo) & (tgt.edge_index[1] < hi)
    remote_nnz = int((own & ((tgt.edge_index[0] // rpr) != rank)).sum()) if distributed else 0
    local_nnz = int(own.sum()) + (hi - lo)
    del own
    torch.manual_seed(0)
    model = DistGRADE(group=group, **hp) if distributed else GRADE(**hp)
    model.grade = model.init_model()
    opt = Adam(model.grade.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    E_ = hp["epoch"]
    import numpy as np
    step = lambda i: model.train_step(src, tgt, 2. / (1. + np.exp(-10. * (i % E_) / E_)) - 1, opt)[0]   # noqa: E731
    ms, loss = _timed_steps(step, warmup, steps, dev, distributed)
    tm = _aggregation_profile(step, 2, lambda m: m[0] == hi - lo and m[1] == H)
    us = statistics.mean(tm) * 1e3 if tm else float("nan")
    b_alg = 4 * (hi - lo + 1) + 8 * local_nnz + 2 * 4 * (hi - lo) * H
    peak, src_ = _peak_hbm()
    # per step and GPU: 2 domains x 2 layers x (forward + backward) aggregations; layer widths H (hidden) and C (5 -> the
    # classifier is a Linear in GRADE, not a conv), so all 8 launches gather H-wide rows
    nvlink_per_launch = remote_nnz * 4 * H
    out = {"config": 4, "workload": "GRADE (Gaussian-MMD) synthetic %dM nodes / %dM edges / %d feat, hid %d, fp32, 1-D node "
                                    "partition over %d x B200, RANDOM graph (no partition locality) "
                                    "(BASELINE.json configs[3])" % (nodes // 10**6, edges // 10**6, F_, H, world),
           "metric": "grade_train_epochs_per_sec", "value": 1e3 / ms, "unit": "epochs/s", "ms_per_step": ms, "n_gpus": world,
           "steps": steps, "warmup": warmup, "dtype": "f32", "scaling": "strong (one graph pair split over the GPUs)",
           "loss": float(loss),
           "partition": {"rows_per_gpu": hi - lo, "local_nnz_target_graph": local_nnz,
                         "remote_column_fraction": remote_frac, "remote_nnz_per_launch_this_gpu": remote_nnz,
                         "nvlink_bytes_per_aggregation_launch_per_gpu": nvlink_per_launch,
                         "nvlink_bytes_per_step_per_gpu_estimate": 8 * nvlink_per_launch,
                         "nvlink_floor_ms_per_step_at_900GBps": 8 * nvlink_per_launch / 900e9 * 1e3,
                         "comm_bound_fraction_of_step": min(1.0, 8 * nvlink_per_launch / 900e9 * 1e3 / ms)},
           "roofline": {"bound": "hbm" if not distributed else "nvlink (remote gathers) / hbm (local)",
                        "kernel": "aggregation at H=%d on this GPU's row block (target graph)" % H,
                        "alg_bytes_per_launch": b_alg, "us_per_launch": us, "achieved": b_alg / us / 1e3, "peak": peak,
                        "unit": "GB/s", "frac": b_alg / us / 1e3 / peak, "peak_source": src_,
                        "launches_profiled_per_step": len(tm) / 2.0},
           "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    del model, opt, src, tgt
    clear_graph_cache()
    torch.cuda.empty_cache()
    return out


def run_config5(dev, rank, world, group, steps=20, warmup=3, graphs=50_000, batch=512):
    """BASELINE.json configs[4]: AdaGCN graph-level, Mutagenicity -> PROTEINS-shaped synthetic datasets of 50k graphs
    per domain (source ~Poisson(30) nodes, 2.05 directed edges per node; target ~Poisson(39), 3.7), 14 one-hot features,
    2 classes, DataLoader(batch_size=512, shuffle=True), hyper-parameters of benchmark/graph/run_all_M.sh:2; data
    parallel over `world` GPUs (every mini-batch's graphs split over the ranks)."""
    import torch
    import torch.distributed as dist
    from pygda_b200.data import DataLoader, DeviceGraphDataset
    from pygda_b200.dist import shard_batch
    from pygda_b200.models import AdaGCN
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import graph_dataset
    distributed = world > 1
    hp = dict(in_dim=14, hid_dim=128, num_classes=2, mode="graph", num_layers=2, dropout=0.4, gnn_type="gcn", adv_dim=40,
              gp_weight=5.0, domain_weight=0.1, weight_decay=0.01, lr=0.01, epoch=400, device=str(dev), batch_size=batch,
              verbose=0)
    t0 = time.perf_counter()
    ds_s = DeviceGraphDataset(graph_dataset(graphs, 30, 2.05, 14, 2, seed=50), dev)
    ds_t = DeviceGraphDataset(graph_dataset(graphs, 39, 3.7, 14, 2, seed=51), dev)
    t_gen = time.perf_counter() - t0
    torch.manual_seed(0)
    if distributed:
        from pygda_b200.models.dist_adagcn import DistAdaGCN
        model = DistAdaGCN(pg=group.pg, **hp)
    else:
        model = AdaGCN(**hp)
    model.adagcn = model.init_model()
    opt = Adam(model.adagcn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    model.init_critic()
    torch.manual_seed(1234)                                   # every rank iterates the same shuffled batches
    loaders = (DataLoader(ds_s, batch_size=batch, shuffle=True, device=dev),
               DataLoader(ds_t, batch_size=batch, shuffle=True, device=dev))
    it = [None]

    def batches():
        while True:
            for sb, tb in zip(*loaders):
                yield sb, tb

    def step(i):
        if it[0] is None:
            it[0] = batches()
        sb, tb = next(it[0])
        if distributed:
            sb, tb = shard_batch(sb, rank, world), shard_batch(tb, rank, world)
        return model.train_step(sb, tb, opt)[0]
    from pygda_b200._lib import load
    lib = load()
    ms_warm, _ = _timed_steps(step, warmup, 1, dev, distributed)
    n0 = lib.gda_launch_count()
    host0 = time.perf_counter()
    ms, loss = _timed_steps(step, 0, steps, dev, distributed)
    host_ms = (time.perf_counter() - host0) * 1e3 / steps
    launches = (lib.gda_launch_count() - n0) / steps
    out = {"config": 5, "workload": "AdaGCN graph-level synthetic Mutagenicity->PROTEINS-shape, %dk graphs per domain, batch "
                                    "%d, %d x B200 data parallel (BASELINE.json configs[4])" % (graphs // 1000, batch, world),
           "metric": "adagcn_train_steps_per_sec", "value": 1e3 / ms, "unit": "steps/s (one step = one mini-batch of %d "
           "graphs per domain: 10 critic iterations + 1 encoder update)" % batch, "ms_per_step": ms, "n_gpus": world,
           "graphs_per_sec": 2 * batch * 1e3 / ms, "epochs_per_sec": 1e3 / ms / ((graphs + batch - 1) // batch),
           "steps": steps, "warmup": warmup, "dtype": "f32", "scaling": "strong (each mini-batch split over the GPUs)",
           "loss": float(loss), "gda_launches_per_step": launches, "host_ms_per_step": host_ms,
           "critic": "closed-form WGAN-GP on libgda (ops.linear / ops.matmul), libgda Adam" if model.analytic_critic
           else "torch double backward", "dataset_seconds": t_gen,
           "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    del model, opt, ds_s, ds_t, loaders
    torch.cuda.empty_cache()
    return out


def run_other_config(args, rank, world, local_rank):
    """`--config 3|4|5`: one JSON line for that BASELINE configuration (builder-run; the driver times config 2)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from pygda_b200.dist import PeerGroup
        group = PeerGroup(device=dev)
    if args.config == 3:
        out = run_config3(dev, max(args.steps, 1), max(args.warmup, 1)) if rank == 0 else None
    elif args.config == 4:
        out = run_config4(dev, rank, world, group, max(args.steps, 1), max(args.warmup, 1))
    else:
        out = run_config5(dev, rank, world, group, max(args.steps, 1), max(args.warmup, 1))
    if rank == 0 and out is not None:
        out.update({"higher_is_better": True, "data": "synthetic", "vs_baseline": None})
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200,
                    help="timed steps (default 200: a timed region above one second; the driver passes its own count)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configuration: 2 = the metric's (default, what the driver times); 3 = UDAGCN 1M/10M "
                         "bf16 on one GPU; 4 = GRADE-MMD 5M/50M over --gpus GPUs (random partition); 5 = AdaGCN graph-level "
                         "50k graphs / batch 512 over --gpus GPUs")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the side measurements of configs 3 / 4 / 5 the default run appends (other_configs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-scale", type=int, default=1,
                    help="tests only: run the reference arm on a 1/SCALE graph pair (the line says so); default full scale")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: device-resident phase only")
    ap.add_argument("--no-cuda-graph", action="store_true", help="issue every kernel from Python (no graph replay)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.config != 2:
        run_other_config(args, rank, world, local_rank)
        return
    run_gpu_arm(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
