"""The body of the full-batch training loop (pygda/models/a2gnn.py:309-319: forward_model,
zero_grad, backward, optimizer.step) captured ONCE as a CUDA graph and replayed.

In full-batch node mode every epoch runs the same ~130 kernels on the same buffers; issuing them
from Python costs about as much host time as the GPU needs to execute them.  Everything that
changes between epochs lives in device memory the graph reads at replay time:

* MMD sample indices: still drawn on the CPU global generator exactly as the reference does
  (pygda/utils/mmd.py:148-149, bit-exact), staged through pinned memory into a fixed device tensor
  before each replay;
* dropout masks: the counter-based seeds are baked, a device-side offset (ops.DropoutRng.offset) is
  advanced by a kernel inside the graph, so every replay draws fresh masks;
* Adam's step count / bias corrections: device state advanced by k_adam_tick;
* the GRL coefficient alpha (a2gnn.py:305-306): a device scalar written before each replay.

Same kernels, same order, same results as the eager step (tests/test_gpu_graphed.py).
"""
import torch

from .. import ops
from .._lib import load
from ..utils import mmd as mmd_utils


class GraphedStep:
    def __init__(self, est, source_data, target_data, optimizer, warmup=2, sampling_num=1000, times=5,
                 alpha_fn=None):
        """``warmup`` eager steps run first (they fill the graph / operand-split / workspace caches that
        the capture must not allocate).  They are REAL optimiser steps -- step i uses ``alpha_fn(i)`` --
        and their (loss, source_logits, target_logits) are kept in ``self.warmup_results``."""
        if getattr(est, "mode", "node") != "node":
            raise ValueError("GraphedStep captures the full-batch node-level step only")
        dev = torch.device(est.device)
        self.est, self.opt = est, optimizer
        self.src, self.tgt = source_data.to(dev), target_data.to(dev)
        self.ns, self.nt = self.src.x.shape[0], self.tgt.x.shape[0]
        self.sampling_num, self.times = sampling_num, times
        self.s_idx = torch.zeros(times, sampling_num, dtype=torch.int64, device=dev)
        self.t_idx = torch.zeros(times, sampling_num, dtype=torch.int64, device=dev)
        self.alpha = torch.zeros(1, dtype=torch.float32, device=dev)
        self.uses_mmd = not getattr(est, "adv", False)
        if ops.dropout_rng.offset is None or ops.dropout_rng.offset.device != dev:
            ops.dropout_rng.enable_device_offset(dev)
        lib = load()

        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                # eager warm-up: fills the graph / split / workspace caches
            self.warmup_results = []
            for i in range(max(int(warmup), 1)):
                self._stage(alpha_fn(i) if alpha_fn is not None else 0.0, None)
                self.warmup_results.append(self._body())
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)

        optimizer.zero_grad(set_to_none=True)        # gradients are re-created inside the graph's memory pool
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.gda_launch_count()
        with torch.cuda.graph(self.graph):
            self.loss, self.source_logits, self.target_logits = self._body()
        self.launches_per_replay = int(lib.gda_launch_count() - n0)
        self.replays = 0

    def _body(self):
        est = self.est
        est.a2gnn.train()
        idx = (self.s_idx, self.t_idx) if self.uses_mmd else None
        loss, s_logits, t_logits = est.forward_model(self.src, self.tgt, self.alpha if not self.uses_mmd else 0.0,
                                                     mmd_indices=idx)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        ops.dropout_rng.advance_offset()
        return loss, s_logits, t_logits

    def _stage(self, alpha, indices):
        if self.uses_mmd:
            if indices is None:
                indices = mmd_utils.draw_indices(self.ns, self.nt, self.sampling_num, self.times)
            mmd_utils.stage_into(self.s_idx, indices[0])
            mmd_utils.stage_into(self.t_idx, indices[1])
        else:
            ops.gda.fill_f32(ops._p(self.alpha), 1, float(alpha), ops._stream())

    def __call__(self, alpha=0.0, mmd_indices=None):
        """One training step; returns (loss, source_logits, target_logits) -- static tensors that the
        next call overwrites."""
        self._stage(alpha, mmd_indices)
        self.graph.replay()
        self.replays += 1
        return self.loss, self.source_logits, self.target_logits
