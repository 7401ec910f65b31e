"""Node-partitioned multi-GPU path (one process per GPU, NVLink peer memory + NCCL).

SURVEY.md section 8e: the graph and every [N, .] activation are split into P contiguous row
blocks.  Rank r keeps CSR rows of its block with columns encoded as (owner, local row); the
aggregation kernel gathers neighbour rows directly from the owners' HBM over NVLink
(``gda_spmm_peer_f32``), a device-side flag barrier (``gda_peer_barrier``) orders the k chained
propagation steps, weight gradients are all-reduced with NCCL once per step, the cross-entropy
mean becomes a scalar all-reduce and the MMD sample rows (5 x 1000 per domain) are exchanged
with one small all-reduce.  The reference has no distributed code at all.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import ops
from ._lib import gda, load
from .data import Data
from .graph import Graph, NORM_SYM_COL, SELF_LOOPS

MAX_PEERS = 8
# Above this share of remote columns the aggregation switches from in-kernel NVLink GATHERS (only referenced rows cross
# the fabric, one 512-byte remote load per non-zero: right for partitions with locality, e.g. a 5 % edge cut) to PUSH
# mode (every output row is stored into every rank's local gather buffer by the kernel that produces it; gathers are
# local).  Measured at BASELINE config 4 on 2 GPUs (random graph, 50 % remote): 31 ms per aggregation with remote
# gathers (~206 GB/s of useful NVLink traffic) -- profiles/r2_f_multi.
PUSH_MIN_REMOTE = 0.25
FUSED_HALO = os.environ.get("GDA_HALO_FUSED", "1") != "0"     # A/B switch: 0 = separate halo push kernel after every step


def _stream():
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def rows_per_rank(n, world):
    """Rows of a contiguous 1-D partition: ceil(n / world); the last rank may own fewer."""
    return (n + world - 1) // world


def block_range(n, world, rank):
    rpr = rows_per_rank(n, world)
    lo = min(n, rank * rpr)
    return lo, min(n, lo + rpr)


def pack_column(col, rpr):
    """(owner << 28 | local row): the column encoding of a partition (gda_graph_partition)."""
    owner = col // rpr
    return (owner << 28) | (col - owner * rpr)


class _DevPtr:
    """Exposes a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class SymBuffer:
    """A cudaMalloc'ed, CUDA-IPC-shared buffer: ``local`` tensor view + the same buffer of every rank."""

    def __init__(self, group, nbytes):
        self.group, self.nbytes = group, int(nbytes)
        ptr = C.c_void_p(0)
        handle = (C.c_ubyte * 64)()
        gda.sym_alloc(self.nbytes, C.byref(ptr), handle)
        self.ptr = ptr.value
        handles = [None] * group.world
        dist.all_gather_object(handles, bytes(handle), group=group.pg)
        self.peer_ptrs, self._opened = [], []
        for q, h in enumerate(handles):
            if q == group.rank:
                self.peer_ptrs.append(self.ptr)
            else:
                p = C.c_void_p(0)
                gda.sym_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(p))
                self.peer_ptrs.append(p.value)
                self._opened.append(p.value)
        self.ptr_array = (C.c_void_p * group.world)(*self.peer_ptrs)
        self.bytes_view = torch.as_tensor(_DevPtr(self.ptr, self.nbytes), device=group.device)

    def view(self, dtype, shape):
        n = 1
        for s in shape:
            n *= s
        return self.bytes_view[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)

    def close(self):
        lib = load()
        for p in self._opened:
            lib.gda_sym_close(C.c_void_p(p))
        self._opened = []
        if self.ptr:
            torch.cuda.synchronize()
            lib.gda_sym_free(C.c_void_p(self.ptr))
            self.ptr = 0


class PeerGroup:
    """The ranks of one NVSwitch box that share a partitioned graph."""

    def __init__(self, device=None, pg=None):
        if not dist.is_initialized():
            raise RuntimeError("init torch.distributed (backend 'nccl') before creating a PeerGroup")
        self.pg = pg
        self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)
        if self.world > MAX_PEERS:
            raise ValueError(f"the peer path supports up to {MAX_PEERS} GPUs of one box")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.flags = SymBuffer(self, 8 * MAX_PEERS)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0
        dist.barrier(group=pg)

    def barrier(self):
        """Device-side all-ranks barrier on the current stream (no host sync)."""
        self.epoch += 1
        gda.peer_barrier(self.flags.ptr_array, self.rank, self.world, self.epoch,
                         C.c_void_p(self.error.data_ptr()), _stream())

    def check(self):
        if int(self.error.item()):
            raise RuntimeError("gda_peer_barrier timed out: a peer rank did not arrive")

    def rows_per_rank(self, n):
        return rows_per_rank(n, self.world)

    def block(self, n):
        return block_range(n, self.world, self.rank)


class PartitionedGraph:
    """Rank-local row block of a normalised graph + the symmetric ping-pong feature buffers."""

    def __init__(self, group, edge_index, num_nodes_global, flags=SELF_LOOPS | NORM_SYM_COL, edge_weight=None):
        self.group = group
        self.global_nodes = int(num_nodes_global)
        self.rows_per_rank = group.rows_per_rank(self.global_nodes)
        self.row_lo, self.row_hi = group.block(self.global_nodes)
        self.num_nodes = self.row_hi - self.row_lo
        # share of this rank's non-zeros whose column lives on another GPU: decides how the exchange is done
        owner = edge_index[0] // self.rows_per_rank
        mine = (edge_index[1] >= self.row_lo) & (edge_index[1] < self.row_hi)
        n_mine = int(mine.sum())
        self.remote_fraction = float((mine & (owner != group.rank)).sum()) / max(n_mine + self.num_nodes, 1)
        del owner, mine
        frac = torch.tensor([self.remote_fraction], device=group.device)
        dist.all_reduce(frac, op=dist.ReduceOp.MAX, group=group.pg)         # every rank must take the same path
        forced = os.environ.get("GDA_DIST_MODE")
        self.mode = forced if forced in ("push", "peer", "halo") else \
            ("push" if float(frac.item()) > PUSH_MIN_REMOTE else "halo")
        self.push = self.mode == "push"
        full = Graph(edge_index, self.global_nodes, edge_weight, flags)      # normalisation needs global degrees

        self._h = C.c_void_p(0)
        gda.graph_partition(full.handle, self.row_lo, self.row_hi, self.rows_per_rank, _stream(), C.byref(self._h))
        n_, nnz_, a_, b_ = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        gda.graph_info(self._h, C.byref(n_), C.byref(nnz_), C.byref(a_), C.byref(b_))
        self.local_nnz = nnz_.value
        self.device = group.device
        self._ws, self._sym = {}, {}
        # every partitioned graph has its own barrier channel (flag array + epoch counter): graphs used
        # on different streams (source / target branch) then never wait on each other's epochs
        self.flags = SymBuffer(group, 8 * MAX_PEERS)
        # the channel's barrier epoch lives in device memory and is advanced by the barrier kernel itself, so that
        # a captured CUDA graph of the training step replays correctly (gda_peer_barrier_dev)
        self.epoch_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._halo = _HaloPlan(self, full) if self.mode == "halo" else None
        del full

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                load().gda_graph_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    def workspace(self, transpose, width):
        key = (int(bool(transpose)), int(width))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = load().gda_spmm_workspace_bytes(self._h, key[0], key[1])
            ws = self._ws[key] = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=self.device)
        return ws

    def _buffers(self, width):
        s = self._sym.get(width)
        if s is None:
            nbytes = self.rows_per_rank * width * 4
            s = self._sym[width] = (SymBuffer(self.group, nbytes), SymBuffer(self.group, nbytes))
        return s

    def _gather_buffers(self, width):
        """Push mode: two symmetric buffers holding the WHOLE (padded) matrix, [world * rows_per_rank, width] each."""
        s = self._sym.get(("push", width))
        if s is None:
            nbytes = self.group.world * self.rows_per_rank * width * 4
            s = self._sym[("push", width)] = (SymBuffer(self.group, nbytes), SymBuffer(self.group, nbytes))
        return s

    def _barrier(self):
        gda.peer_barrier_dev(self.flags.ptr_array, self.group.rank, self.group.world, ops._p(self.epoch_dev),
                             C.c_void_p(self.group.error.data_ptr()), _stream())

    def supports_nb(self, h, nb):
        """True when ``nb`` stacked matrices can be aggregated in one pass (halo mode, H = 128)."""
        return nb == 1 or (nb == 2 and self._halo is not None and h == 128)

    def spmm_k(self, x, k, transpose=False, bias=None, relu=False, dropout_p=0.0, seed=0, seed_offset=None, nb=1):
        """A_hat^k x over the partition: x and the result are this rank's [n_local, H] blocks (``nb`` of them
        stacked: the paired bottleneck evaluations, halo mode only)."""
        x = ops._f32c(x)
        rows, h = x.shape
        n = rows // nb
        if n != self.num_nodes or n * nb != rows:
            raise ValueError(f"x has {rows} rows, this rank owns {self.num_nodes} (x {nb})")
        if not self.supports_nb(h, nb):
            raise NotImplementedError("stacked matrices are aggregated in one pass in halo mode at H = 128 only")
        g = self.group
        ws = self.workspace(transpose, h)
        out = torch.empty(rows, h, dtype=torch.float32, device=self.device)
        flags = (ops.EPI_RELU if relu else 0) | (ops.EPI_DROPOUT if dropout_p > 0 else 0)
        if self.push and h == 128:
            return self._spmm_k_push(x, k, transpose, bias, flags, dropout_p, seed, seed_offset, ws, out)
        if self._halo is not None and h == 128:
            return self._halo.spmm_k(x, k, transpose, bias, flags, dropout_p, seed, seed_offset, out, nb)
        bufs = self._buffers(h)
        if ops.PROFILE is None:
            # barrier + copy-in + k x (barrier, peer aggregation) behind ONE call: k+2 fewer host round trips
            gda.spmm_peer_k_dev_f32(self._h, int(bool(transpose)), int(k), ops._p(x), bufs[0].ptr_array,
                                    bufs[1].ptr_array, g.world, g.rank, ops._p(out), h, ops._p(bias), flags,
                                    float(dropout_p), int(seed) & 0xFFFFFFFFFFFFFFFF, ops._p(seed_offset), ops._p(ws),
                                    ws.numel(), self.flags.ptr_array, ops._p(self.epoch_dev),
                                    C.c_void_p(g.error.data_ptr()), _stream())
            return out
        # bench.py instrumentation: one call per step so that each launch can be timed
        views = [b.view(torch.float32, (self.rows_per_rank, h)) for b in bufs]
        self._barrier()                               # every rank is done with the buffers of the last call
        views[0][:n].copy_(x)
        for i in range(k):
            last = i == k - 1
            self._barrier()                           # step i-1 (or the copy-in) is complete on every rank
            dst = out if last else views[(i + 1) & 1]
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            gda.spmm_peer_f32(self._h, int(bool(transpose)), bufs[i & 1].ptr_array, g.world, g.rank, h,
                              ops._p(dst), h, h, ops._p(bias if last else None), flags if last else 0,
                              float(dropout_p if last else 0.0), int(seed) & 0xFFFFFFFFFFFFFFFF,
                              ops._p(seed_offset), ops._p(ws), ws.numel(), _stream())
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            ops.PROFILE.append((e0, e1, (n, h, "float32", 1, "weighted-peer")))
        return out


    def _spmm_k_push(self, x, k, transpose, bias, flags, dropout_p, seed, seed_offset, ws, out):
        """A_hat^k x with the exchange fused into the producing kernel (gda_spmm_push_k_f32)."""
        g, (n, h) = self.group, x.shape
        b0, b1 = self._gather_buffers(h)
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        err = C.c_void_p(g.error.data_ptr())
        if ops.PROFILE is None:
            gda.spmm_push_k_f32(self._h, int(bool(transpose)), int(k), ops._p(x), b0.ptr_array, b1.ptr_array, g.world,
                                g.rank, ops._p(out), h, ops._p(bias), flags, float(dropout_p), seed, ops._p(seed_offset),
                                ops._p(ws), ws.numel(), self.flags.ptr_array, ops._p(self.epoch_dev), err, _stream())
            return out
        # bench.py instrumentation: the same sequence, one call per step so that each launch can be timed
        block = self.rows_per_rank * h * 4
        self._barrier()
        for q in range(g.world):
            dst = torch.as_tensor(_DevPtr(b0.peer_ptrs[q] + g.rank * block, n * h * 4), device=self.device)
            dst.view(torch.float32).view(n, h).copy_(x)
        for i in range(k):
            last = i == k - 1
            self._barrier()
            src, dstb = (b1, b0) if i & 1 else (b0, b1)
            outs = [out.data_ptr()] if last else [dstb.peer_ptrs[q] + g.rank * block for q in range(g.world)]
            arr = (C.c_void_p * len(outs))(*outs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gda.spmm_push_f32(self._h, int(bool(transpose)), C.c_void_p(src.peer_ptrs[g.rank]), arr, len(outs), g.world,
                              h, h, h, ops._p(bias if last else None), flags if last else 0,
                              float(dropout_p if last else 0.0), seed, ops._p(seed_offset), ops._p(ws), ws.numel(),
                              _stream())
            e1.record()
            ops.PROFILE.append((e0, e1, (n, h, "float32", 1, "weighted-push" if not last else "weighted-push-last")))
        return out


class _HaloPlan:
    """HALO mode of a partitioned graph (few remote columns): this rank's row block as a RECTANGULAR local graph --
    columns [0, n_own) are its own rows, columns n_own.. are halo slots, one per distinct remote row it references --
    plus the list of its own rows that other ranks reference.  An aggregation then runs on local memory with the
    ordinary work-list kernel (entries of a row keep the global COO order, so results are those of the single-GPU
    kernel); after every step each rank copies the rows its peers need into their halo slots (gda_push_rows_f32, P2P
    stores over NVLink): a referenced row crosses the fabric once per consumer and step, not once per non-zero."""

    def __init__(self, part, full):
        from .graph import SKIP_EMPTY_ROWS
        g, dev = part.group, part.device
        self.part = part
        lo, hi, rpr = part.row_lo, part.row_hi, part.rows_per_rank
        n_own = hi - lo
        ei, w = full.coo()                                   # normalised weights, the reference's entry order
        src, dst = ei[0], ei[1]

        def block(rows, cols):
            m = (rows >= lo) & (rows < hi)
            return rows[m] - lo, cols[m], w[m]
        fr, fc, fw = block(dst, src)                         # A_hat   : row = target, column = source
        tr, tc, tw = block(src, dst)                         # A_hat^T : row = source, column = target
        remote = lambda c: c[(c < lo) | (c >= hi)]           # noqa: E731
        halo = torch.unique(torch.cat([remote(fc), remote(tc)]))          # sorted global ids of the referenced remote rows
        self.n_own, self.n_halo = n_own, int(halo.numel())
        self.n_tot = n_own + self.n_halo

        def remap(c):
            own = (c >= lo) & (c < hi)
            slot = torch.searchsorted(halo, c.clamp(min=0)) if self.n_halo else torch.zeros_like(c)
            return torch.where(own, c - lo, n_own + slot)
        self.graphs = (Graph(torch.stack([remap(fc), fr]), self.n_tot, fw, SKIP_EMPTY_ROWS),
                       Graph(torch.stack([remap(tc), tr]), self.n_tot, tw, SKIP_EMPTY_ROWS))
        # who needs which of my rows, and where it goes in their buffer
        owner = (halo // rpr).cpu()
        local_row = (halo - (halo // rpr) * rpr).cpu()
        slots = torch.arange(n_own, self.n_tot)
        need = {q: (local_row[owner == q].to(torch.int32), slots[owner == q].to(torch.int32)) for q in range(g.world)
                if q != g.rank}
        everyone = [None] * g.world
        dist.all_gather_object(everyone, need, group=g.pg)
        rows, slot, peer = [], [], []
        for p_, wanted in enumerate(everyone):
            if p_ == g.rank or g.rank not in wanted:
                continue
            r_, s_ = wanted[g.rank]
            rows.append(r_); slot.append(s_); peer.append(torch.full((r_.numel(),), p_, dtype=torch.int32))
        cat = lambda xs: (torch.cat(xs) if xs else torch.zeros(0, dtype=torch.int32)).to(dev)   # noqa: E731
        self.push_rows, self.push_slots, self.push_peer = cat(rows), cat(slot), cat(peer)
        self.push_count = int(self.push_rows.numel())
        # the same plan per row, for the kernel that pushes while it aggregates (gda_spmm_halo_f32)
        self.halo_mask = torch.zeros(max(n_own, 1), dtype=torch.int32, device=dev)
        self.halo_slot = torch.zeros(max(n_own, 1) * g.world, dtype=torch.int32, device=dev)
        if self.push_count:
            r64, p64 = self.push_rows.long(), self.push_peer.long()
            self.halo_mask.index_put_((r64,), torch.bitwise_left_shift(torch.ones_like(self.push_peer), self.push_peer),
                                      accumulate=True)                       # (row, peer) pairs are distinct: add == or
            self.halo_slot[r64 * g.world + p64] = self.push_slots
        self._bufs = {}
        # The factored chain (D S (D^2 S)^(k-1) D, no per-edge weights: the single-GPU default) on the block: the
        # block's weights are dinv[row] * dinv[col] of the WHOLE unit-weight graph, halo rows arrive scaled by their
        # owners, so only this rank's own dinv entries are needed (gda_graph_set_unit_dinv).
        self.unit = False
        if load().gda_graph_unit_weights(full.handle, 0, 128, 1) == 1 and os.environ.get("GDA_HALO_UNW", "1") != "0":
            dinv_full = torch.empty(part.global_nodes, dtype=torch.float32, device=dev)
            gda.graph_export_dinv(full.handle, ops._p(dinv_full), _stream())
            dinv_blk = torch.zeros(max(self.n_tot, 1), dtype=torch.float32, device=dev)
            dinv_blk[:n_own] = dinv_full[lo:hi]
            for gr in self.graphs:
                gda.graph_set_unit_dinv(gr.handle, ops._p(dinv_blk), _stream())
            torch.cuda.current_stream(dev).synchronize()           # dinv_blk / dinv_full may be freed now
            self.unit = all(load().gda_graph_unit_weights(gr.handle, 0, 128, nb_) == 1
                            for gr in self.graphs for nb_ in (1, 2))
        # two stacked matrices: matrix m lives at rows [m * n_tot, (m + 1) * n_tot) of a rank's buffers -- n_tot differs
        # from rank to rank, so a pushed row of matrix 1 lands n_tot OF THE RECEIVER rows further down
        all_tot = [None] * g.world
        dist.all_gather_object(all_tot, self.n_tot, group=g.pg)
        self.peer_n_tot = [int(v) for v in all_tot]
        self.push2 = None
        if self.push_count:
            peer_tot = torch.tensor(self.peer_n_tot, dtype=torch.int32, device=dev)[self.push_peer.long()]
            self.push2 = (torch.cat([self.push_rows, self.push_rows + self.n_tot]),
                          torch.cat([self.push_slots, self.push_slots + peer_tot]),
                          torch.cat([self.push_peer, self.push_peer]))

    def buffers(self, width, nb=1):
        b = self._bufs.get((width, nb))
        if b is None:
            nbytes = nb * max(self.n_tot, 1) * width * 4
            syms = (SymBuffer(self.part.group, nbytes), SymBuffer(self.part.group, nbytes))
            b = self._bufs[(width, nb)] = (syms, tuple(s_.view(torch.float32, (nb * self.n_tot, width)) for s_ in syms))
        return b

    def _push(self, view, sym, width, nb=1):
        if self.push_count:
            g = self.part.group
            rows, slots, peer = (self.push_rows, self.push_slots, self.push_peer) if nb == 1 else self.push2
            gda.push_rows_f32(ops._p(view), width, ops._p(rows), ops._p(slots), ops._p(peer), nb * self.push_count,
                              sym.ptr_array, g.world, width, width, _stream())

    def spmm_k(self, x, k, transpose, bias, flags, dropout_p, seed, seed_offset, out, nb=1):
        part, h = self.part, x.shape[1]
        n, nt = x.shape[0] // nb, self.n_tot
        gr = self.graphs[1 if transpose else 0]
        syms, views = self.buffers(h, nb)
        ws = gr.workspace(False, h * nb)
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        unit = self.unit
        part._barrier()                                   # every rank is done with the buffers of the last call
        if unit:                                          # z0 = D x on this rank's own rows
            gda.row_scale_rows_f32(gr.handle, n, nb, ops._p(x), h, n * h, ops._p(views[0]), h, nt * h, h, _stream())
        else:
            for m in range(nb):
                views[0][m * nt:m * nt + n].copy_(x[m * n:(m + 1) * n])
        self._push(views[0], syms[0], h, nb)
        for i in range(k):
            last = i == k - 1
            part._barrier()                               # step i-1 (or the copy-in) and its halo pushes have landed
            src, dst_v, dst_s = views[i & 1], views[(i + 1) & 1], syms[(i + 1) & 1]
            prof = ops.PROFILE
            if prof is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if unit or nb > 1:
                y, ys = (out, n * h) if last else (dst_v, nt * h)
                tail = (ops._p(bias if last else None), flags if last else 0, float(dropout_p if last else 0.0), seed,
                        ops._p(seed_offset), ops._p(ws), ws.numel(), _stream())
                fused = unit and not last and FUSED_HALO
                if fused:                                 # z' = D^2 S z, halo rows stored to their consumers as they finish
                    strides = (C.c_int64 * part.group.world)(*[t_ * h for t_ in self.peer_n_tot])
                    gda.spmm_unw_halo_f32(gr.handle, nb, ops._p(src), h, nt * h, ops._p(y), h, ys, h,
                                          ops._p(self.halo_mask), ops._p(self.halo_slot), dst_s.ptr_array, strides,
                                          part.group.world, ops._p(ws), ws.numel(), _stream())
                elif unit:                                # (y = D S z on the last step): no edge weights
                    gda.spmm_unw_nb_f32(gr.handle, 0, nb, ops._p(src), h, nt * h, ops._p(y), h, ys, h, int(last), *tail)
                else:
                    gda.spmm_nb_f32(gr.handle, 0, nb, ops._p(src), h, nt * h, ops._p(y), h, ys, h, *tail)
                if prof is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    prof.append((e0, e1, (n, h, "float32", nb, "unit-weight-halo" if unit else "weighted-halo")))
                if not last and not fused:
                    self._push(dst_v, dst_s, h, nb)
                continue
            if last or not FUSED_HALO:
                gda.spmm_f32(gr.handle, 0, ops._p(src), h, ops._p(out if last else dst_v), h, h,
                             ops._p(bias if last else None), flags if last else 0, float(dropout_p if last else 0.0),
                             seed, ops._p(seed_offset), ops._p(ws), ws.numel(), _stream())
            else:                                         # aggregate and push the halo rows in one kernel
                gda.spmm_halo_f32(gr.handle, ops._p(src), h, ops._p(dst_v), h, h, ops._p(self.halo_mask),
                                  ops._p(self.halo_slot), dst_s.ptr_array, part.group.world, None, 0, 0.0, seed,
                                  ops._p(seed_offset), ops._p(ws), ws.numel(), _stream())
            if prof is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                prof.append((e0, e1, (n, h, "float32", 1, "weighted-halo")))
            if not last and not FUSED_HALO:
                self._push(dst_v, dst_s, h)
        return out


class _PartitionTag:
    """Attached to a partitioned ``edge_index`` so that conv layers pick the peer path."""

    def __init__(self, group, num_nodes_global):
        self.group, self.num_nodes_global = group, num_nodes_global
        self._graphs = {}

    def graph(self, edge_index, flags, edge_weight=None):
        g = self._graphs.get(flags)
        if g is None:
            g = self._graphs[flags] = PartitionedGraph(self.group, edge_index, self.num_nodes_global, flags,
                                                       edge_weight)
        return g


def attach_partition(data, group):
    """Tag a rank-local ``Data`` (x / y rows of this rank's block + the GLOBAL edge_index, with
    ``num_nodes_global`` / ``row_lo`` / ``row_hi`` set) so that conv layers take the peer path."""
    lo, hi = group.block(data.num_nodes_global)
    if (lo, hi) != (data.row_lo, data.row_hi):
        raise ValueError("data block does not match this rank's row range")
    data.edge_index._gda_partition = _PartitionTag(group, data.num_nodes_global)
    if torch.is_tensor(data.x) and data.x.is_cuda:
        data.x._gda_const = True                # resident block of input features (ops.ConstCache)
    return data


def partition_data(data, group):
    """Rank-local view of a full graph: x / y rows of this rank's block, the whole edge_index
    (tagged), on the group's device."""
    n = data.x.size(0)
    lo, hi = group.block(n)
    ei = data.edge_index.to(group.device)
    ei._gda_partition = _PartitionTag(group, n)
    out = Data(x=data.x[lo:hi].to(group.device), edge_index=ei, y=data.y[lo:hi].to(group.device))
    out.x._gda_const = True
    out.num_nodes_global, out.row_lo, out.row_hi = n, lo, hi
    return out


# ------------------------------------------------------------------ collectives with autograd
class AllReduceSum(torch.autograd.Function):
    """sum over ranks of a (scaled) local term; backward is the identity on the local term."""

    @staticmethod
    def forward(ctx, t, pg):
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=pg)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


class GatherRows(torch.autograd.Function):
    """rows ``idx`` (global ids) of a row-partitioned matrix, replicated on every rank: each rank
    fills the rows it owns and one all-reduce completes them.  Backward: every rank keeps the
    gradient of the rows it owns (scatter-add, indices repeat)."""

    @staticmethod
    def forward(ctx, feats, idx, row_lo, pg):
        n = feats.size(0)
        local = idx - row_lo
        own = ((local >= 0) & (local < n)).to(feats.dtype).unsqueeze(1)
        safe = local.clamp(0, max(n - 1, 0))          # no nonzero(): that would sync the host every step
        out = feats.index_select(0, safe) * own
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=pg)
        ctx.save_for_backward(safe, own)
        ctx.shape = feats.shape
        return out

    @staticmethod
    def backward(ctx, g):
        safe, own = ctx.saved_tensors
        gx = torch.zeros(ctx.shape, dtype=g.dtype, device=g.device)
        gx.index_add_(0, safe, g * own)
        return gx, None, None, None


class AllGatherRows(torch.autograd.Function):
    """Row blocks of every rank concatenated in rank order, replicated on every rank (blocks may differ in
    length).  Backward: a rank keeps the gradient of its own block -- correct when what is computed from the
    gathered matrix is REPLICATED on every rank (every rank holds the same upstream gradient), as the pooled
    graph encodings of the data-parallel graph-level path are (models/dist_adagcn.py)."""

    @staticmethod
    def forward(ctx, x, pg):
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
        counts = torch.zeros(world, dtype=torch.int64, device=x.device)
        counts[rank] = x.size(0)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=pg)
        counts = [int(c) for c in counts.tolist()]
        out = torch.zeros(sum(counts), *x.shape[1:], dtype=x.dtype, device=x.device)
        lo = sum(counts[:rank])
        out[lo:lo + x.size(0)] = x
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=pg)      # blocks are disjoint: sum == concatenation
        ctx.lo, ctx.n = lo, x.size(0)
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.lo:ctx.lo + ctx.n].contiguous(), None


def shard_batch(batch, rank, world):
    """Graphs [lo, hi) of a collated ``Batch`` (contiguous 1-D split of its graphs over the ranks) as a new
    ``Batch`` with local node ids -- the data-parallel split of one mini-batch."""
    from .data import Batch
    g = len(batch)
    lo, hi = block_range(g, world, rank)
    ptr = getattr(batch, "ptr", None)
    if ptr is None:
        counts = torch.bincount(batch.batch, minlength=g)
        ptr = torch.zeros(g + 1, dtype=torch.long, device=counts.device)
        ptr[1:] = torch.cumsum(counts, 0)
    n_lo, n_hi = int(ptr[lo]), int(ptr[hi])
    ei = batch.edge_index
    keep = (ei[1] >= n_lo) & (ei[1] < n_hi)
    return Batch(x=batch.x[n_lo:n_hi], edge_index=(ei[:, keep] - n_lo).contiguous(), y=batch.y[lo:hi],
                 batch=batch.batch[n_lo:n_hi] - lo, num_graphs=hi - lo, ptr=(ptr[lo:hi + 1] - n_lo).clone())


def allreduce_grads(params, pg=None):
    """Sum the weight gradients over ranks (one flat NCCL all-reduce)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=pg)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
