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


class StagedBatch:
    """Per-epoch host->device staging of ONE pinned host batch into the FIXED device tensors a captured step reads.

    The reference's loop copies the batch to the device at the top of every step (pygda/models/a2gnn.py:311-312) and
    only then computes.  Here the copy of epoch e+1 is issued on a copy stream into one of two device staging slots
    while epoch e's graph runs; at the top of epoch e+1 the slot is consumed on the compute stream: a row-compressed
    ``x`` is rebuilt densely into the static ``x`` buffer (``gda_unpack_tiles_f32``, bit for bit), the other tensors
    are copied device-to-device, and the cached split-bf16 operand pair of ``x`` is recomputed from it
    (``ops.ConstCache.refresh``) -- every epoch's data crosses PCIe and is consumed; nothing is re-allocated.

    When the first layer multiplies from the TILE-PACKED form itself (``ops.x_tiles``: sparse bag-of-words
    features, csrc/gemm_xt.cu) the epoch's packed arrays are copied device-to-device into the static packed
    buffers the captured kernels read, and that is all: no kernel of the step reads the dense ``x`` (it keeps epoch
    0's rebuild for consumers outside the step, e.g. ``predict``), so neither the dense rebuild nor the operand
    split is repeated."""

    def __init__(self, host_batch, device):
        self.dev = torch.device(device)
        self.static = host_batch.to(self.dev)            # epoch 0's copy; allocates the static device tensors
        self.packed = host_batch.__dict__.get("_packed_x")
        self.fields = {}
        if self.packed is not None:
            self.fields.update(self.packed.tensors())
        for k, v in host_batch.__dict__.items():
            if torch.is_tensor(v) and not v.is_cuda and not (k == "x" and self.packed is not None):
                if k == "edge_index" and v.dtype == torch.int64 and v.numel() and int(v.max()) < 2 ** 31 \
                        and int(v.min()) >= 0:
                    # node ids cross PCIe as int32 (half the bytes) and are widened on the device; like the packed x,
                    # this staging form is built once, when the loader's batch is first staged
                    self.fields[k] = v.to(torch.int32).pin_memory()
                else:
                    self.fields[k] = v if v.is_pinned() else v.pin_memory()
        self.tiles = ops.x_tiles(self.static.x, 128) if self.packed is not None else None
        if self.tiles is not None and self.tiles is not getattr(self.static.x, "_gda_tiles", None):
            self.tiles = None
        self.slots = [{k: torch.empty(v.shape, dtype=v.dtype, device=self.dev) for k, v in self.fields.items()}
                      for _ in range(2)]
        self.copied, self.consumed = [None, None], [None, None]
        self.stream = torch.cuda.Stream(self.dev)
        self.nbytes = sum(v.numel() * v.element_size() for v in self.fields.values())

    def issue(self, slot):
        """Queue the host->device copy of the whole batch into ``slot`` on the copy stream."""
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])
            for k, v in self.fields.items():
                self.slots[slot][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.copied[slot] = ev

    def consume(self, slot):
        """On the current stream: wait for the slot's copy and move it into the static tensors."""
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self.copied[slot])
        s = self.slots[slot]
        if self.tiles is not None:
            for k, t in self.tiles.tensors().items():
                if k in s:
                    t.copy_(s[k], non_blocking=True)
            if "_vals" not in s:                          # exponent-packed staging: rebuild the fp32 values in place
                self.packed.unpack_values_into(s, self.tiles.vals)
        elif self.packed is not None:
            self.packed.unpack_into(s, self.static.x)
        for k, v in s.items():
            if not k.startswith("_"):
                getattr(self.static, k).copy_(v, non_blocking=True)
        if self.tiles is None:
            ops.split_cache.refresh(self.static.x)
            ops.bf16_cache.refresh(self.static.x)
        ev = torch.cuda.Event()
        ev.record(main)
        self.consumed[slot] = ev


def _on_device(data, dev):
    return all((not torch.is_tensor(v)) or v.device == dev for v in data.__dict__.values())


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
        # host-resident batches are re-sent every epoch like the reference does (a2gnn.py:311-312), double-buffered
        self.staged = [StagedBatch(d, dev) if not _on_device(d, dev) else None for d in (source_data, target_data)]
        self.src = self.staged[0].static if self.staged[0] is not None else source_data
        self.tgt = self.staged[1].static if self.staged[1] is not None else target_data
        self.h2d_bytes_per_step = sum(sb.nbytes for sb in self.staged if sb is not None)
        self._slot = 0
        # MMD sample indices range over the WHOLE graph (a rank of a partitioned run holds a row block of it)
        self.ns = int(getattr(self.src, "num_nodes_global", self.src.x.shape[0]))
        self.nt = int(getattr(self.tgt, "num_nodes_global", self.tgt.x.shape[0]))
        self.group = getattr(est, "group", None)         # partitioned estimator (models/dist_a2gnn.py)
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
        for sb in self.staged:                           # the first replay's inputs start crossing PCIe now
            if sb is not None:
                sb.issue(self._slot)

    def _body(self):
        est = self.est
        est.a2gnn.train()
        idx = (self.s_idx, self.t_idx) if self.uses_mmd else None
        loss, s_logits, t_logits = est.forward_model(self.src, self.tgt, self.alpha if not self.uses_mmd else 0.0,
                                                     mmd_indices=idx)
        est.backward_and_step(loss, self.opt)            # zero_grad, backward, [gradient all-reduce,] Adam step
        ops.dropout_rng.advance_offset()
        return loss, s_logits, t_logits

    def _stage(self, alpha, indices):
        if self.uses_mmd:
            if indices is None:
                indices = mmd_utils.draw_indices(self.ns, self.nt, self.sampling_num, self.times)
            mmd_utils.stage_into(self.s_idx, indices[0])
            mmd_utils.stage_into(self.t_idx, indices[1])
            if self.group is not None and self.group.world > 1:
                # one draw for all ranks: rank 0's (pygda/utils/mmd.py:148-149), as DistA2GNN._mmd_indices does
                import torch.distributed as dist
                dist.broadcast(self.s_idx, src=0, group=self.group.pg)
                dist.broadcast(self.t_idx, src=0, group=self.group.pg)
        else:
            ops.gda.fill_f32(ops._p(self.alpha), 1, float(alpha), ops._stream())

    def _restage(self, prefetch_next=True):
        """Consume this epoch's staged inputs, then start the next epoch's copy (it overlaps the replay)."""
        n0 = load().gda_launch_count()
        for sb in self.staged:
            if sb is not None:
                sb.consume(self._slot)
        self.restage_launches = int(load().gda_launch_count() - n0)
        self._slot ^= 1
        if prefetch_next:
            for sb in self.staged:
                if sb is not None:
                    sb.issue(self._slot)

    def __call__(self, alpha=0.0, mmd_indices=None, last=False):
        """One training step; returns (loss, source_logits, target_logits) -- static tensors that the
        next call overwrites.  ``last``: no further step follows (do not start another host->device copy)."""
        self._restage(prefetch_next=not last)
        self._stage(alpha, mmd_indices)
        self.graph.replay()
        self.replays += 1
        return self.loss, self.source_logits, self.target_logits
