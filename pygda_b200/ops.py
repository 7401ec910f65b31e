"""torch.autograd.Function shims over the C ABI (include/gda.h).

Everything numeric on the hot path is a libgda kernel; torch supplies device
memory, the autograd tape and the current stream.  Backward passes run on
PyTorch's autograd thread: every call therefore reads the *current* stream of
the calling thread.
"""
import ctypes as C

import torch

from ._lib import gda, load

EPI_RELU, EPI_DROPOUT = 1, 2
_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice
_NULL = C.c_void_p(0)
# bench.py sets this to a list to collect (start_event, end_event, (N, H, dtype)) per aggregation launch
PROFILE = None


def _stream():
    """The calling thread's CURRENT CUDA stream as a raw handle.  torch.cuda.current_stream() builds a Stream object
    through several Python layers (16 us per call -- 9 ms of the 47 ms host-bound AdaGCN mini-batch step,
    profiles/r2_l_config5_host); the raw accessor is the same query without them."""
    return C.c_void_p(_raw_stream(_cur_device()))


def _p(t):
    return _NULL if t is None else C.c_void_p(t.data_ptr())


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError("pygda_b200 kernels run on the GPU only (no CPU fallback)")
    return t if t.is_contiguous() else t.contiguous()


def _bf16c(t):
    if t.dtype != torch.bfloat16:
        raise TypeError(f"expected bfloat16, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError("pygda_b200 kernels run on the GPU only (no CPU fallback)")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------ raw kernels
_GOLDEN = 0x9E3779B97F4A7C15
_M64 = 0xFFFFFFFFFFFFFFFF


def shift_seed(seed, elements):
    """Seed under which mask index i equals mask index ``elements + i`` of ``seed`` (common.cuh mix_hash)."""
    return (int(seed) + int(elements) * _GOLDEN) & _M64


def spmm(graph, x, transpose=False, bias=None, relu=False, dropout_p=0.0, seed=0, out=None,
         seed_offset=None, nb=1):
    """One aggregation Y = A_hat X (or A_hat^T X) with the optional fused epilogue.
    ``nb`` > 1: x is ``nb`` stacked [N, H] matrices that share the graph."""
    x = x if x.is_contiguous() else x.contiguous()
    rows, h = x.shape
    n = rows // nb
    if n != graph.num_nodes or n * nb != rows:
        raise ValueError(f"x has {rows} rows but the graph has {graph.num_nodes} nodes (x {nb})")
    if out is None:
        out = torch.empty_like(x)
    ws = graph.workspace(transpose, h * nb)
    flags = (EPI_RELU if relu else 0) | (EPI_DROPOUT if dropout_p > 0 else 0)
    tail = (h, _p(bias), flags, float(dropout_p), int(seed) & _M64, _p(seed_offset), _p(ws), ws.numel(), _stream())
    prof = PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    if nb > 1:
        if x.dtype != torch.float32:
            raise TypeError("batched aggregation supports float32 features")
        gda.spmm_nb_f32(graph.handle, int(bool(transpose)), nb, _p(x), h, n * h, _p(out), h, n * h, *tail)
    elif x.dtype == torch.float32:
        gda.spmm_f32(graph.handle, int(bool(transpose)), _p(x), h, _p(out), h, *tail)
    elif x.dtype == torch.bfloat16:
        gda.spmm_bf16(graph.handle, int(bool(transpose)), _p(x), h, _p(out), h, *tail)
    else:
        raise TypeError(f"aggregation supports float32 and bfloat16 features, got {x.dtype}")
    if prof is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        prof.append((e0, e1, (n, h, str(x.dtype).replace("torch.", ""), nb, "weighted")))
    return out


def unit_weight_chain(graph, transpose, h, nb, k):
    """True when A_hat^k on this graph runs as the factored chain D S (D^2 S)^(k-1) D (gda_spmm_unw_nb_f32)."""
    return k >= 3 and not hasattr(graph, "spmm_k") and \
        load().gda_graph_unit_weights(graph.handle, int(bool(transpose)), int(h), int(nb)) == 1


def _spmm_k_unw_profiled(graph, x, k, transpose, bias, relu, dropout_p, seed, seed_offset, nb):
    """bench.py's instrumented form of the factored chain: the same calls gda_spmm_k_nb_f32 makes, one ABI call
    per step so that each launch can be bracketed by CUDA events."""
    rows, h = x.shape
    n = rows // nb
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    out = torch.empty_like(x)
    ws = graph.workspace(transpose, h * nb)
    gda.row_scale_f32(graph.handle, nb, _p(x), h, n * h, _p(bufs[0]), h, n * h, h, _stream())
    flags = (EPI_RELU if relu else 0) | (EPI_DROPOUT if dropout_p > 0 else 0)
    src = bufs[0]
    for i in range(k):
        last = i == k - 1
        dst = out if last else bufs[(i + 1) & 1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gda.spmm_unw_nb_f32(graph.handle, int(bool(transpose)), nb, _p(src), h, n * h, _p(dst), h, n * h, h, int(last),
                            _p(bias) if last else _NULL, flags if last else 0, float(dropout_p) if last else 0.0,
                            int(seed) & _M64, _p(seed_offset), _p(ws), ws.numel(), _stream())
        e1.record()
        PROFILE.append((e0, e1, (n, h, "float32", nb, "unit-weight")))
        src = dst
    return out


def spmm_k(graph, x, k, transpose=False, bias=None, relu=False, dropout_p=0.0, seed=0,
           seed_offset=None, nb=1):
    """A_hat^k X with ping-pong buffers; the epilogue is applied on the last step."""
    if k == 0:
        raise ValueError("spmm_k needs k >= 1")
    if hasattr(graph, "spmm_k"):          # row-partitioned graph: NVLink path (pygda_b200/dist.py)
        return graph.spmm_k(x, k, transpose=transpose, bias=bias, relu=relu, dropout_p=dropout_p, seed=seed,
                            seed_offset=seed_offset, nb=nb)
    if x.dtype == torch.float32 and PROFILE is not None and x.is_contiguous() and \
            unit_weight_chain(graph, transpose, x.shape[1], nb, k):
        return _spmm_k_unw_profiled(graph, x, k, transpose, bias, relu, dropout_p, seed, seed_offset, nb)
    if x.dtype != torch.float32 or PROFILE is not None:
        # bf16 features, or bench.py timing each launch: one ABI call per step
        cur, bufs = x, [None, None]
        for i in range(k):
            last = i == k - 1
            dst = bufs[i & 1]
            if dst is None:
                dst = bufs[i & 1] = torch.empty_like(x)
            cur = spmm(graph, cur, transpose, bias if last else None, relu and last,
                       dropout_p if last else 0.0, seed, out=dst, seed_offset=seed_offset, nb=nb)
        return cur
    # fp32: all k launches behind one call (gda_spmm_k_nb_f32) -- k-fold less host work per conv
    x = x if x.is_contiguous() else x.contiguous()
    rows, h = x.shape
    n = rows // nb
    if n != graph.num_nodes or n * nb != rows:
        raise ValueError(f"x has {rows} rows but the graph has {graph.num_nodes} nodes (x {nb})")
    out = torch.empty_like(x)
    t0 = torch.empty_like(x) if k >= 2 else None
    t1 = torch.empty_like(x) if k >= 3 else None
    ws = graph.workspace(transpose, h * nb)
    flags = (EPI_RELU if relu else 0) | (EPI_DROPOUT if dropout_p > 0 else 0)
    gda.spmm_k_nb_f32(graph.handle, int(bool(transpose)), int(k), int(nb), _p(x), h, n * h, _p(out), h, n * h, _p(t0),
                      _p(t1), h, _p(bias), flags, float(dropout_p), int(seed) & _M64, _p(seed_offset), _p(ws),
                      ws.numel(), _stream())
    return out


def gemm(a, b, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, out=None):
    """out[M,N] = alpha * op(a) @ op(b) + beta * out   (fp32, row-major)."""
    a, b = _f32c(a), _f32c(b)
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    kb, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if k != kb:
        raise ValueError(f"gemm inner dimensions differ: {k} vs {kb}")
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    nbytes = load().gda_gemm_workspace_bytes(int(trans_a), int(trans_b), m, n, k)
    ws = _workspace(nbytes, a.device)
    gda.gemm_f32(int(trans_a), int(trans_b), m, n, k, float(alpha), _p(a), a.stride(0), _p(b),
                 b.stride(0), float(beta), _p(out), out.stride(0), _p(ws), ws.numel(), _stream())
    return out


class Split:
    """fp32 matrix as a split-bf16 pair (hi + lo) for the tcgen05 GEMM; ``ld`` in elements."""
    __slots__ = ("hi", "lo", "rows", "cols", "ld")

    def __init__(self, x):
        x = _f32c(x)
        self.rows, self.cols = x.shape
        self.ld = (self.cols + 7) // 8 * 8
        self.hi = torch.empty(self.rows, self.ld, dtype=torch.bfloat16, device=x.device)
        self.lo = torch.empty(self.rows, self.ld, dtype=torch.bfloat16, device=x.device)
        self.fill(x)

    def fill(self, x):
        """(Re)compute the pair from ``x`` into the existing buffers."""
        x = _f32c(x)
        gda.split_bf16(_p(x), self.rows, self.cols, x.stride(0), _p(self.hi), _p(self.lo), self.ld, _stream())

    @property
    def nbytes(self):
        return 2 * self.hi.numel() * 2


class ConstCache:
    """Derived forms (split-bf16 pair, bf16 copy) of CONSTANT operands -- the input features ``data.x``.

    Key: the HOST tensor a device copy was made from when ``Data.to`` attached one (``_gda_key``, as for
    ``edge_index``): the fit loops re-send the same host graph every step (pygda/models/a2gnn.py:311-312), each time
    to a fresh device tensor -- such a copy maps to ONE entry whose buffers are refilled in place from the new copy
    (the step's data is consumed, nothing is re-allocated, nothing stale is kept alive).  Tensors without a host key
    are keyed by (data_ptr, version) and kept alive so that the pointer cannot be recycled.
    Only tensors MARKED constant are inserted -- a host key from ``Data.to``, or ``mark_constant`` (the loaders mark
    the device-resident ``data.x`` they hand out): "does not require grad" is not enough, under ``torch.no_grad()``
    (predict) every activation looks like that and would evict the real constants.  Anything else is computed
    uncached.  Bounded by entries and by bytes."""

    def __init__(self, make, fill, nbytes, capacity=8, max_bytes=48 << 30):
        self._make, self._fill, self._nbytes = make, fill, nbytes
        self.capacity, self.max_bytes, self._d = capacity, max_bytes, {}

    @staticmethod
    def _key(x):
        k = getattr(x, "_gda_key", None)
        if k is not None:
            return ("host", k, tuple(x.shape), str(x.device))
        return ("dev", x.data_ptr(), x._version, tuple(x.shape), str(x.device))

    def get(self, x):
        key = self._key(x)
        hit = self._d.get(key)
        if hit is not None:
            if hit[2] != x.data_ptr():                   # a new device copy of the same host tensor: refill in place
                self._fill(hit[0], x)
                hit[2] = x.data_ptr()
            return hit[0]
        val = self._make(x)
        if key[0] != "host" and not getattr(x, "_gda_const", False):
            return val
        while self._d and (len(self._d) >= self.capacity or
                           sum(self._nbytes(v[0]) for v in self._d.values()) + self._nbytes(val) > self.max_bytes):
            self._d.pop(next(iter(self._d)))
        # host-keyed entries keep the HOST tensor alive (its address is the key), not a stale device copy
        keep = getattr(x, "_gda_keepalive", x) if key[0] == "host" else x
        self._d[key] = [val, keep, x.data_ptr()]
        return val

    def refresh(self, x):
        """``x`` was re-written in place (a staged copy landed in its buffer): recompute the cached form."""
        hit = self._d.get(self._key(x))
        if hit is not None:
            self._fill(hit[0], x)
            hit[2] = x.data_ptr()
        return hit is not None

    def clear(self):
        self._d.clear()


def mark_constant(t):
    """Declare ``t`` (a device tensor that is not re-written between steps, e.g. the input features of a full-batch
    loader) eligible for the operand caches."""
    if torch.is_tensor(t):
        t._gda_const = True
    return t


def SplitCache(capacity=8):
    return ConstCache(Split, lambda sp, x: sp.fill(x), lambda sp: sp.nbytes, capacity)


split_cache = SplitCache()
TC_MIN_FLOPS = 5e8      # below this the SIMT kernel is memory-bound anyway


def tc_eligible(m, n, k):
    return n >= 64 and k >= 64 and 2.0 * m * n * k >= TC_MIN_FLOPS and \
        load().gda_gemm_bf16x3_supported(m, n, k, 8, 8) == 1


def gemm_split(a, b, trans_a, trans_b, m, n, k):
    """C[m,n] = op(a) op(b) from Split operands on the tcgen05 kernel."""
    out = torch.empty(m, n, dtype=torch.float32, device=a.hi.device)
    nbytes = load().gda_gemm_bf16x3_workspace_bytes(m, n, k)
    ws = _workspace(nbytes, out.device)
    gda.gemm_bf16x3(int(trans_a), int(trans_b), m, n, k, _p(a.hi), _p(a.lo), a.ld, _p(b.hi), _p(b.lo), b.ld,
                    _p(out), out.stride(0), _p(ws), ws.numel(), _stream())
    return out


# ---- first-layer products from a TILE-PACKED sparse input matrix (csrc/gemm_xt.cu) ----
XT_DENSITY = 0.10        # above this the sub-tiles outgrow the expanders' register prefetch: dense split path
XT_MIN_COLS = 512        # fewer than 8 K blocks per CTA: the pipeline never fills, the dense kernel streams as fast
XT_ENABLED = True


def _make_tiles(x):
    """Tile-packed form of a device-resident CONSTANT dense matrix, or False when it is not sparse enough."""
    from .data import PackedTiles
    x = _f32c(x)
    nnz = int((x.view(torch.int32) != 0).sum())
    if nnz > XT_DENSITY * x.numel():
        return False
    return PackedTiles(x, pin=False).view()


tiles_cache = ConstCache(_make_tiles, lambda t, x: None, lambda t: t.nbytes if t else 0, capacity=8)


def x_tiles(x, n_out):
    """``XTiles`` to multiply ``x`` [rows, cols] from, or None: the packed copy ``Data.to`` attached to a staged
    ``x`` (``_gda_tiles``), or -- for a device-resident matrix marked constant (``mark_constant`` / the full-batch
    loaders) -- a packed form built once (``tiles_cache``) when at most ``XT_DENSITY`` of it is non-zero."""
    if not XT_ENABLED or x.dim() != 2 or x.dtype != torch.float32 or x.requires_grad or x.shape[1] < XT_MIN_COLS:
        return None
    if not tc_eligible(x.shape[0], max(int(n_out), 64), x.shape[1]):
        return None
    t = getattr(x, "_gda_tiles", None)
    if t is not None and t.dense_version is not None and t.dense_version != x._version:
        t = None                               # the dense matrix was edited in place after the packed copy was made
    if t is None and getattr(x, "_gda_const", False) and getattr(x, "_gda_key", None) is None:
        t = tiles_cache.get(x)
    if not t or t.vals.numel() > XT_DENSITY * x.numel():
        return None
    return t


def _prof_begin():
    """bench.py's instrumented repeat (``PROFILE`` is a list): a CUDA event before the launch."""
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(e0, meta):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((e0, e1, meta))


def _xt_args(t):
    return _p(t.vals), _p(t.codes), _p(t.ptr), _p(t.seg), t.rows, t.cols


def xt_fwd(t, w, w_in_out):
    """x @ w (``w_in_out``: w [cols, N]) or x @ w^T (w [N, cols]) from the tile-packed x."""
    ws = Split(w)
    n = w.shape[1] if w_in_out else w.shape[0]
    out = torch.empty(t.rows, n, dtype=torch.float32, device=w.device)
    ev = _prof_begin()
    gda.gemm_xt_fwd(*_xt_args(t), int(not w_in_out), n, _p(ws.hi), _p(ws.lo), ws.ld, _p(out), out.stride(0),
                    _stream())
    _prof_end(ev, (t.rows, t.cols, "xt_fwd", n, int(t.vals.numel())))
    return out, ws


def xt_dw(t, g, w_in_out):
    """Weight gradient of ``xt_fwd``: x^T g [cols, N] (``w_in_out``) or g^T x [N, cols]; returns (dW, split of g)."""
    g = _f32c(g)
    gs = Split(g)
    n = g.shape[1]
    out = torch.empty((t.cols, n) if w_in_out else (n, t.cols), dtype=torch.float32, device=g.device)
    ws = _workspace(load().gda_gemm_xt_dw_workspace_bytes(t.rows, t.cols, n), g.device)
    ev = _prof_begin()
    gda.gemm_xt_dw(*_xt_args(t), n, _p(gs.hi), _p(gs.lo), gs.ld, int(not w_in_out), _p(out), out.stride(0), _p(ws),
                   ws.numel(), _stream())
    _prof_end(ev, (t.rows, t.cols, "xt_dw", n, int(t.vals.numel())))
    return out, gs


def _pad_to(t, dim, size):
    """Zero-padded copy of a 2-D tensor along ``dim``."""
    shape = list(t.shape)
    shape[dim] = size
    out = torch.zeros(shape, dtype=t.dtype, device=t.device)
    out.narrow(dim, 0, t.shape[dim]).copy_(t)
    return out


TC_PAD = 64      # the tcgen05 kernel needs N >= 64 and K >= 64


def mm(a, b, trans_a=False, trans_b=False, a_split=None, b_split=None, cache_a=False, cache_b=False):
    """op(a) @ op(b): tensor cores (split-bf16, fp32-accurate) for the big shapes, SIMT otherwise.
    Returns (result, a_split, b_split) so callers can re-use the splits in the backward.

    Tall products with a NARROW output or reduction width between the skinny kernels' 16 and the tensor-core
    kernel's 64 -- UDAGCN's / AdaGCN's 40-wide domain MLP on 1 M rows (pygda/nn/udagcn_base.py:157-162) -- are
    zero-padded to 64 and run on the tensor cores too: exact (the padding contributes zeros), and 10x faster than
    the SIMT kernel there (profiles/r1_h_config3: 4 x 1.9 ms of a 28 ms step)."""
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    n = b.shape[0] if trans_b else b.shape[1]
    if tc_eligible(m, n, k):
        if a_split is None:
            a_split = split_cache.get(a) if cache_a else Split(a)
        if b_split is None:
            b_split = split_cache.get(b) if cache_b else Split(b)
        return gemm_split(a_split, b_split, trans_a, trans_b, m, n, k), a_split, b_split
    if 16 < n < TC_PAD and tc_eligible(m, TC_PAD, k):                 # narrow output: pad op(b)'s columns
        bp = Split(_pad_to(_f32c(b), 0 if trans_b else 1, TC_PAD))
        ap = a_split if a_split is not None else (split_cache.get(a) if cache_a else Split(a))
        out = gemm_split(ap, bp, trans_a, trans_b, m, TC_PAD, k)
        return out[:, :n].contiguous(), ap, b_split
    if 16 < k < TC_PAD and tc_eligible(m, n, TC_PAD):                 # narrow reduction: pad both operands along k
        ap = Split(_pad_to(_f32c(a), 0 if trans_a else 1, TC_PAD))
        bp = Split(_pad_to(_f32c(b), 1 if trans_b else 0, TC_PAD))
        return gemm_split(ap, bp, trans_a, trans_b, m, n, TC_PAD), a_split, b_split
    return gemm(a, b, trans_a=trans_a, trans_b=trans_b), a_split, b_split


def colsum(x):
    x = _f32c(x)
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    gda.colsum_f32(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), _stream())
    return out


def scale(x, alpha):
    x = _f32c(x)
    y = torch.empty_like(x)
    gda.scale_f32(_p(y), _p(x), x.numel(), float(alpha), _stream())
    return y


class LinCombFn(torch.autograd.Function):
    """y = wa * a + wb * b on libgda (gda_scale_f32 + gda_axpy_f32): the sum of two branch outputs
    (pygda/nn/mixup_gcnconv.py:212) and mixup's interpolation ``a * lam + b * (1 - lam)``
    (pygda/nn/mixup_base.py:141,148,156)."""

    @staticmethod
    def forward(ctx, a, b, wa, wb):
        a, b = _f32c(a), _f32c(b)
        if a.shape != b.shape:
            raise ValueError(f"shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
        y = torch.empty_like(a)
        gda.scale_f32(_p(y), _p(a), a.numel(), float(wa), _stream())
        gda.axpy_f32(_p(y), _p(b), b.numel(), float(wb), _stream())
        ctx.w = (float(wa), float(wb))
        return y

    @staticmethod
    def backward(ctx, g):
        wa, wb = ctx.w
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = g if wa == 1.0 else scale(g, wa)
        if ctx.needs_input_grad[1]:
            gb = g if wb == 1.0 else scale(g, wb)
        return ga, gb, None, None


def add(a, b):
    return LinCombFn.apply(a, b, 1.0, 1.0)


def lerp2(a, b, lam):
    """a * lam + b * (1 - lam)."""
    return LinCombFn.apply(a, b, float(lam), 1.0 - float(lam))


class DropoutRng:
    """Seeds for the counter-based dropout masks.

    The reference draws dropout masks from the CUDA generator and the MMD sample
    indices from the CPU generator (pygda/utils/mmd.py:148-149); to keep the latter
    bit-exact nothing here touches the CPU generator.  seed = f(torch.cuda seed,
    call counter); ``offset`` is an optional device uint64 added inside the kernels so
    that a captured CUDA graph gets a fresh mask per replay (gda_counter_inc)."""

    def __init__(self):
        self.counter = 0
        self.offset = None

    def next(self):
        self.counter += 1
        base = torch.cuda.initial_seed() if torch.cuda.is_available() else torch.initial_seed()
        return (base * 0x9E3779B97F4A7C15 + self.counter * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF

    def enable_device_offset(self, device):
        self.offset = torch.zeros(1, dtype=torch.int64, device=device)
        return self.offset

    def advance_offset(self):
        if self.offset is not None:
            gda.counter_inc(_p(self.offset), _stream())


dropout_rng = DropoutRng()


def next_seed():
    return dropout_rng.next()


# ------------------------------------------------------------------ autograd nodes
class PropagateFn(torch.autograd.Function):
    """A_hat^k x; backward (A_hat^T)^k g.  (prop_gcn_conv.py:208-210 / cached_gcn_conv.py:138)"""

    @staticmethod
    def forward(ctx, x, graph, k):
        ctx.graph, ctx.k = graph, k
        return spmm_k(graph, x, k)

    @staticmethod
    def backward(ctx, g):
        return spmm_k(ctx.graph, g.contiguous(), ctx.k, transpose=True), None, None


class GraphConvFn(torch.autograd.Function):
    """y = A_hat^k (x W^T) + b as ONE tape node (PropGCNConv.forward,
    prop_gcn_conv.py:205-213; CachedGCNConv.forward with ``w_in_out``,
    cached_gcn_conv.py:130-138,171-172).  k = 0 is the plain Linear + bias."""

    @staticmethod
    def forward(ctx, x, weight, bias, graph, k, w_in_out):
        x = _f32c(x)
        w = _f32c(weight)
        const_x = not x.requires_grad          # input features: split once, re-used every pass
        tiles = x_tiles(x, w.shape[1] if w_in_out else w.shape[0])
        if tiles is not None:                  # sparse input features: multiplied from their tile-packed form
            (h, ws), xs = xt_fwd(tiles, w, w_in_out), None
        else:
            h, xs, ws = mm(x, w, trans_b=not w_in_out, cache_a=const_x)
        if k > 0:
            y = spmm_k(graph, h, k, bias=bias)
        elif bias is not None:
            y = h
            gda.bias_act_dropout_fwd(_p(h), _p(bias), _p(y), h.shape[0], h.shape[1], 0, 0.0, 0, _NULL, _stream())
        else:
            y = h
        ctx.save_for_backward(x, w)
        ctx.graph, ctx.k, ctx.w_in_out, ctx.has_bias = graph, k, w_in_out, bias is not None
        ctx.splits = (xs, ws)
        ctx.tiles = tiles
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        xs, ws = ctx.splits
        gy = _f32c(gy)
        gb = colsum(gy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        g0 = spmm_k(ctx.graph, gy, ctx.k, transpose=True) if ctx.k > 0 else gy
        gw = gx = gs = None
        if ctx.needs_input_grad[1]:
            # W [out,in]: dW = g0^T x ; W [in,out]: dW = x^T g0
            if ctx.tiles is not None:
                gw, gs = xt_dw(ctx.tiles, g0, ctx.w_in_out)
            elif ctx.w_in_out:
                gw, _, gs = mm(x, g0, trans_a=True, a_split=xs)
            else:
                gw, gs, _ = mm(g0, x, trans_a=True, b_split=xs)
        if ctx.needs_input_grad[0]:
            gx, _, _ = mm(g0, w, trans_b=ctx.w_in_out, a_split=gs, b_split=ws)
        return gx, gw, gb, None, None, None


class LinearFn(torch.autograd.Function):
    """nn.Linear: y = x W^T + b (heads at a2gnn_base.py:67,70; udagcn_base.py:155-162)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, w = _f32c(x), _f32c(weight)
        tiles = x_tiles(x, w.shape[0])
        if tiles is not None:
            (y, ws), xs = xt_fwd(tiles, w, False), None
        else:
            y, xs, ws = mm(x, w, trans_b=True, cache_a=not x.requires_grad)
        if bias is not None:
            gda.bias_act_dropout_fwd(_p(y), _p(bias), _p(y), y.shape[0], y.shape[1], 0, 0.0, 0, _NULL, _stream())
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.splits = (xs, ws)
        ctx.tiles = tiles
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        xs, ws = ctx.splits
        gy = _f32c(gy)
        gx = gw = gs = None
        if ctx.needs_input_grad[1]:
            if ctx.tiles is not None:
                gw, gs = xt_dw(ctx.tiles, gy, False)
            else:
                gw, gs, _ = mm(gy, x, trans_a=True, b_split=xs)
        if ctx.needs_input_grad[0]:
            gx, _, _ = mm(gy, w, a_split=gs, b_split=ws)
        gb = colsum(gy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


class ActDropoutFn(torch.autograd.Function):
    """dropout(relu(x)) (or dropout alone) with the keep mask regenerated from the seed."""

    @staticmethod
    def forward(ctx, x, act, p, seed):
        off = dropout_rng.offset if p > 0 else None
        if x.dtype == torch.bfloat16:                     # bf16 feature path (BASELINE config 3)
            x = _bf16c(x)
            y = torch.empty_like(x)
            gda.act_dropout_bf16_fwd(_p(x), _p(y), x.numel(), act, float(p), seed, _p(off), _stream())
            ctx.save_for_backward(y)
            ctx.cfg = (act, float(p), seed, 0, 0, off)
            return y
        x = _f32c(x)
        y = torch.empty_like(x)
        rows = x.shape[0] if x.dim() > 1 else 1
        cols = x.numel() // max(rows, 1)
        gda.bias_act_dropout_fwd(_p(x), _NULL, _p(y), rows, cols, act, float(p), seed, _p(off), _stream())
        ctx.save_for_backward(y)
        ctx.cfg = (act, float(p), seed, rows, cols, off)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        act, p, seed, rows, cols, off = ctx.cfg
        if y.dtype == torch.bfloat16:
            gy = _bf16c(gy)
            gx = torch.empty_like(gy)
            gda.act_dropout_bf16_bwd(_p(gy), _p(y), _p(gx), gy.numel(), act, p, seed, _p(off), _stream())
            return gx, None, None, None
        gy = _f32c(gy)
        gx = torch.empty_like(gy)
        gda.bias_act_dropout_bwd(_p(gy), _p(y), _p(gx), _NULL, rows, cols, act, p, seed, _p(off), _stream())
        return gx, None, None, None


def act_dropout(x, act, p, training):
    """``dropout(act(x))`` as the reference applies it (a2gnn_base.py:136-138).
    ``act`` is a callable; relu is fused, anything else is applied by the caller's
    callable and only the dropout runs here."""
    import torch.nn.functional as F
    p_eff = float(p) if training else 0.0
    if act is F.relu or act is torch.relu:
        return ActDropoutFn.apply(x, 1, p_eff, next_seed() if p_eff > 0 else 0)
    if act is not None:
        x = act(x)
    if p_eff > 0:
        return ActDropoutFn.apply(x, 0, p_eff, next_seed())
    return x


# ------------------------------------------------------------------ paired bottleneck evaluations
# The reference evaluates feat_bottleneck twice per domain and step on the same graph and weights, with fresh
# dropout masks (pygda/models/a2gnn.py:181 & :192, :193 & :211).  Layer 1 is shared (dropout acts after it);
# from there on the two evaluations travel as one stacked [2N, .] matrix: one GEMM, and one aggregation launch
# per propagation step that reads every (colidx, weight) once for both (gda_spmm_nb_f32).  Values are those of
# two separate evaluations with the masks of rows [0, N) and [N, 2N) of the stacked matrix.
def _stack2(a, b):
    """[2N, C] matrix of two [N, C] halves -- a view when they already sit back to back in one buffer."""
    if (a.is_contiguous() and b.is_contiguous() and a.shape == b.shape and a.dtype == b.dtype and
            a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr() and
            b.storage_offset() == a.storage_offset() + a.numel()):
        return a.new_empty(0).set_(a.untyped_storage(), a.storage_offset(), (2 * a.shape[0], a.shape[1]),
                                   (a.shape[1], 1))
    return torch.cat([a, b], 0)


class ActDropoutPairFn(torch.autograd.Function):
    """x -> (dropout_1(act(x)), dropout_2(act(x))): one read, two independently masked copies."""

    @staticmethod
    def forward(ctx, x, act, p, seed):
        x = _f32c(x)
        n, c = x.shape
        y = torch.empty(2 * n, c, dtype=torch.float32, device=x.device)
        off = dropout_rng.offset if p > 0 else None
        gda.bias_act_dropout_rep_fwd(_p(x), _NULL, _p(y), n, c, 2, act, float(p), seed, _p(off), _stream())
        ctx.save_for_backward(y)
        ctx.cfg = (act, float(p), seed, n, c, off)
        ctx.set_materialize_grads(False)
        return y[:n], y[n:]

    @staticmethod
    def backward(ctx, ga, gb):
        if ga is None and gb is None:
            return None, None, None, None
        (y,) = ctx.saved_tensors
        act, p, seed, n, c, off = ctx.cfg
        ga = _f32c(ga) if ga is not None else None
        gb = _f32c(gb) if gb is not None else None
        gx = torch.empty(n, c, dtype=torch.float32, device=y.device)
        gda.bias_act_dropout_rep_bwd(_p(ga), _p(gb), _p(y), _p(gx), _NULL, n, c, 2, 1, act, p, seed, _p(off), _stream())
        return gx, None, None, None


class PairGraphConvActFn(torch.autograd.Function):
    """(xa, xb) -> dropout(act(A_hat^k (x W^T) + b)) for both halves as ONE tape node: stacked GEMM, batched
    aggregation with the bias / ReLU / dropout epilogue fused into its last step (k = 0: one elementwise pass).
    A half whose output receives no gradient (the target-side evaluation that only yields ``target_logits``,
    a2gnn.py:211) costs nothing in the backward."""

    @staticmethod
    def forward(ctx, xa, xb, weight, bias, graph, k, w_in_out, act, p, seed):
        x = _stack2(_f32c(xa), _f32c(xb))
        w = _f32c(weight)
        n = x.shape[0] // 2
        h, _, ws = mm(x, w, trans_b=not w_in_out)
        off = dropout_rng.offset if p > 0 else None
        if k > 0:
            y = spmm_k(graph, h, k, bias=bias, relu=bool(act), dropout_p=p, seed=seed, seed_offset=off, nb=2)
        else:
            y = h
            if bias is not None or act or p > 0:
                gda.bias_act_dropout_fwd(_p(h), _p(bias), _p(y), 2 * n, h.shape[1], int(act), float(p), seed, _p(off),
                                         _stream())
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (graph, k, w_in_out, bias is not None, int(act), float(p), seed, off, n)
        ctx.w_split = ws
        ctx.set_materialize_grads(False)
        return y[:n], y[n:]

    @staticmethod
    def backward(ctx, ga, gb):
        none = (None,) * 10
        if ga is None and gb is None:
            return none
        x, w, y = ctx.saved_tensors
        graph, k, w_in_out, has_bias, act, p, seed, off, n = ctx.cfg
        c = y.shape[1]
        both = ga is not None and gb is not None
        if both:
            nb, xs, ys, sd = 2, x, y, seed
        elif ga is not None:
            nb, xs, ys, sd = 1, x[:n], y[:n], seed
        else:                                              # masks of the second half: rows [N, 2N) of the stack
            nb, xs, ys, sd = 1, x[n:], y[n:], shift_seed(seed, n * c)
        g0_, g1_ = (ga, gb) if both else ((ga if ga is not None else gb), None)
        g0_ = _f32c(g0_)
        g1_ = _f32c(g1_) if g1_ is not None else None
        if act or p > 0 or both:
            g = torch.empty(nb * n, c, dtype=torch.float32, device=y.device)
            gda.bias_act_dropout_rep_bwd(_p(g0_), _p(g1_), _p(ys), _p(g), _NULL, n, c, nb, 0, act, p, sd, _p(off),
                                         _stream())
        else:
            g = g0_
        gbias = colsum(g) if (has_bias and ctx.needs_input_grad[3]) else None
        g0 = spmm_k(graph, g, k, transpose=True, nb=nb) if k > 0 else g
        gw = gx = gs = None
        if ctx.needs_input_grad[2]:
            if w_in_out:
                gw, _, gs = mm(xs, g0, trans_a=True)
            else:
                gw, gs, _ = mm(g0, xs, trans_a=True)
        gxa = gxb = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gx, _, _ = mm(g0, w, trans_b=w_in_out, a_split=gs, b_split=ctx.w_split)
            if both:
                gxa, gxb = gx[:n], gx[n:]
            elif ga is not None:
                gxa = gx
            else:
                gxb = gx
        return (gxa, gxb, gw, gbias) + (None,) * 6


def act_dropout_pair(x, p):
    """Two independently masked copies of dropout(relu(x)), p > 0."""
    return ActDropoutPairFn.apply(x, 1, float(p), next_seed())


def graph_conv_act_pair(xa, xb, weight, bias, graph, k, p, relu=True, w_in_out=False):
    return PairGraphConvActFn.apply(xa, xb, weight, bias, graph, int(k), bool(w_in_out), 1 if relu else 0,
                                    float(p), next_seed() if p > 0 else 0)


# ------------------------------------------------------------------ bf16 feature path (BASELINE config 3)
# "UDAGCN ... hid=256, bf16": input features and activations live in bf16, parameters, accumulation, biases and
# losses in fp32.  The layer GEMMs run as ONE tcgen05 UMMA per K step on bf16 operands (gda_gemm_bf16), the
# aggregation gathers 2-byte rows (gda_spmm_bf16).
def to_bf16(x):
    x = _f32c(x)
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    gda.cast_f32_bf16(_p(x), _p(y), x.numel(), _stream())
    return y


def to_f32(x):
    x = _bf16c(x)
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    gda.cast_bf16_f32(_p(x), _p(y), x.numel(), _stream())
    return y


class CastFn(torch.autograd.Function):
    """bf16 <-> fp32 boundary of the bf16 feature path; the gradient is cast the other way."""

    @staticmethod
    def forward(ctx, x, to_half):
        ctx.to_half = to_half
        return to_bf16(x) if to_half else to_f32(x)

    @staticmethod
    def backward(ctx, g):
        return (to_f32(g) if ctx.to_half else to_bf16(g)), None


def _bf16_fill(y, x):
    x = _f32c(x)
    gda.cast_f32_bf16(_p(x), _p(y), x.numel(), _stream())


def Bf16Cache(capacity=8):
    """bf16 copies of CONSTANT fp32 tensors (the input features), keyed like SplitCache."""
    return ConstCache(to_bf16, _bf16_fill,
                      lambda y: y.numel() * 2, capacity)


bf16_cache = Bf16Cache()


def gemm_bf16(a, b, trans_a=False, trans_b=False, out_bf16=True):
    """op(a) @ op(b) on bf16 operands, fp32 accumulation; bf16 or fp32 result."""
    a, b = _bf16c(a), _bf16c(b)
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    kb, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if k != kb:
        raise ValueError(f"gemm inner dimensions differ: {k} vs {kb}")
    lib = load()
    if lib.gda_gemm_bf16x3_supported(m, n, k, a.stride(0), b.stride(0)) == 1 and \
            a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0:
        out = torch.empty(m, n, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=a.device)
        ws = _workspace(lib.gda_gemm_bf16_workspace_bytes(m, n, k, int(out_bf16)), a.device)
        gda.gemm_bf16(int(trans_a), int(trans_b), m, n, k, _p(a), a.stride(0), _p(b), b.stride(0), _p(out),
                      out.stride(0), int(out_bf16), _p(ws), ws.numel(), _stream())
        return out
    # shapes the tensor-core kernel does not take (tiny widths): the fp32 kernels on the same bf16 VALUES
    out = gemm(to_f32(a), to_f32(b), trans_a=trans_a, trans_b=trans_b)
    return to_bf16(out) if out_bf16 else out


def colsum_bf16(x):
    x = _bf16c(x)
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    gda.colsum_bf16(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), _stream())
    return out


class GraphConvBf16Fn(torch.autograd.Function):
    """GraphConvFn on bf16 rows: y = bf16(A_hat^k (x bf16(W)) + b), fp32 accumulation throughout.  Weight and
    bias gradients are fp32; the gradient w.r.t. x (if needed) is bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, graph, k, w_in_out):
        x = _bf16c(x)
        w16 = to_bf16(weight)
        h = gemm_bf16(x, w16, trans_b=not w_in_out, out_bf16=True)
        if k > 0:
            y = spmm_k(graph, h, k, bias=bias)
        elif bias is not None:
            y = to_bf16(BiasAddFn.forward(None, to_f32(h), bias))
        else:
            y = h
        ctx.save_for_backward(x, w16)
        ctx.graph, ctx.k, ctx.w_in_out, ctx.has_bias = graph, k, w_in_out, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w16 = ctx.saved_tensors
        gy = _bf16c(gy)
        gb = colsum_bf16(gy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        g0 = spmm_k(ctx.graph, gy, ctx.k, transpose=True) if ctx.k > 0 else gy
        gw = gx = None
        if ctx.needs_input_grad[1]:
            if ctx.w_in_out:
                gw = gemm_bf16(x, g0, trans_a=True, out_bf16=False)          # [in, out] = x^T g0
            else:
                gw = gemm_bf16(g0, x, trans_a=True, out_bf16=False)          # [out, in] = g0^T x
        if ctx.needs_input_grad[0]:
            gx = gemm_bf16(g0, w16, trans_b=ctx.w_in_out, out_bf16=True)
        return gx, gw, gb, None, None, None


class BiasAddFn(torch.autograd.Function):
    """x + bias (broadcast over rows) on libgda; backward = (g, column sums of g)."""

    @staticmethod
    def forward(ctx, x, bias):
        x = _f32c(x)
        y = torch.empty_like(x)
        gda.bias_act_dropout_fwd(_p(x), _p(bias), _p(y), x.shape[0], x.shape[1], 0, 0.0, 0, _NULL, _stream())
        return y

    @staticmethod
    def backward(ctx, g):
        return g, colsum(g)


class GradReverse(torch.autograd.Function):
    """pygda/nn/reverse_layer.py:4-66: identity forward, -alpha * g backward."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if torch.is_tensor(ctx.alpha):          # device scalar: a captured CUDA graph reads the current value
            g = _f32c(g)
            out = torch.empty_like(g)
            gda.scale_dev_f32(_p(out), _p(g), g.numel(), -1.0, _p(ctx.alpha), _stream())
            return out, None
        return scale(g, -float(ctx.alpha)), None


def _validate_labels(labels, c):
    """F.nll_loss raises for a target outside [0, C) (pygda/models/a2gnn.py:182); so does this path -- once per label
    tensor OBJECT (and per host tensor it was copied from, ``Data.to`` links the two): the labels of a full-batch fit
    are the same tensor every step, so there is no per-step host synchronisation."""
    src = getattr(labels, "_gda_src", None)
    if getattr(labels, "_gda_labels_ok", None) == c or (src is not None and getattr(src, "_gda_labels_ok", None) == c):
        return
    if torch.cuda.is_current_stream_capturing():
        return                                           # validated by the warm-up step that precedes every capture
    if labels.numel():
        lo, hi = int(labels.min()), int(labels.max())
        if lo < 0 or hi >= c:
            raise IndexError(f"Target {lo if lo < 0 else hi} is out of bounds.")
    labels._gda_labels_ok = c
    if src is not None:
        src._gda_labels_ok = c


class SoftmaxCEFn(torch.autograd.Function):
    """mean CE(log_softmax(logits), labels); labels=None means row r has label (r >= split)."""

    @staticmethod
    def forward(ctx, logits, labels, split):
        z = _f32c(logits)
        rows, c = z.shape
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        if labels is not None:
            labels = labels.contiguous()
            if labels.dtype != torch.int64 or labels.numel() != rows:
                raise ValueError("labels must be int64 with one entry per row")
            _validate_labels(labels, c)
        gda.softmax_ce_fwd_bwd(_p(z), rows, c, z.stride(0), _p(labels), int(split), _p(loss), _p(dz), _stream())
        ctx.save_for_backward(dz)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return _scale_by_scalar(dz, g), None, None


class SoftmaxEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        z = _f32c(logits)
        rows, c = z.shape
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        gda.softmax_entropy_fwd_bwd(_p(z), rows, c, z.stride(0), _p(loss), _p(dz), _stream())
        ctx.save_for_backward(dz)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return _scale_by_scalar(dz, g)


def _scale_by_scalar(t, g):
    """t * g for a 0-dim DEVICE scalar g, without a host sync."""
    out = torch.empty_like(t)
    gs = g.reshape(1).to(torch.float32)
    gda.scale_dev_f32(_p(out), _p(t), t.numel(), 1.0, _p(gs), _stream())
    return out


class CombineFn(torch.autograd.Function):
    """sum_i w_i * t_i over 0-dim device scalars (the loss combination at
    models/a2gnn.py:183-209, udagcn.py:189-199) in one tiny kernel."""

    @staticmethod
    def forward(ctx, weights, *terms):
        n = len(terms)
        ts = [t.reshape(1).to(torch.float32) for t in terms]
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        ws = (C.c_float * n)(*[float(w) for w in weights])
        out = torch.empty((), dtype=torch.float32, device=ts[0].device)
        gda.combine_scalars(n, ptrs, ws, _p(out), _stream())
        ctx.weights = [float(w) for w in weights]
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(1).to(torch.float32).contiguous()
        outs = []
        for w in ctx.weights:
            y = torch.empty((), dtype=torch.float32, device=g.device)
            gda.scale_f32(_p(y), _p(g), 1, w, _stream())
            outs.append(y)
        return (None, *outs)


def combine(pairs):
    """pairs: [(scalar_tensor, weight), ...] -> sum_i weight_i * tensor_i."""
    return CombineFn.apply([w for _, w in pairs], *[t for t, _ in pairs])


class MMDFn(torch.autograd.Function):
    """pygda/utils/mmd.py:109-158 with the sample indices as inputs."""

    @staticmethod
    def forward(ctx, src, tgt, src_idx, tgt_idx, kernel_mul, kernel_num):
        src, tgt = _f32c(src), _f32c(tgt)
        times, b = src_idx.shape
        d = src.shape[1]
        if tgt.shape[1] != d or tgt_idx.shape != src_idx.shape:
            raise ValueError("MMD: source/target feature widths and sample shapes must match")
        nbytes = load().gda_mmd_workspace_bytes(times, b, d)
        ws = _workspace(nbytes, src.device)
        loss = torch.empty((), dtype=torch.float32, device=src.device)
        gda.mmd_fwd(_p(src), src.stride(0), _p(tgt), tgt.stride(0), d, _p(src_idx), _p(tgt_idx), times, b,
                    float(kernel_mul), int(kernel_num), _p(loss), _p(ws), ws.numel(), _stream())
        ctx.save_for_backward(src, tgt, src_idx, tgt_idx, ws)
        return loss

    @staticmethod
    def backward(ctx, g):
        src, tgt, src_idx, tgt_idx, ws = ctx.saved_tensors
        times, b = src_idx.shape
        d = src.shape[1]
        gs, gt = torch.empty_like(src), torch.empty_like(tgt)
        gda.fill_f32(_p(gs), gs.numel(), 0.0, _stream())
        gda.fill_f32(_p(gt), gt.numel(), 0.0, _stream())
        gscale = g.reshape(1).to(torch.float32).contiguous()
        gda.mmd_bwd(_p(src), src.stride(0), _p(tgt), tgt.stride(0), d, _p(src_idx), _p(tgt_idx), times, b,
                    _p(gscale), _p(gs), gs.stride(0), _p(gt), gt.stride(0), _p(ws), ws.numel(), _stream())
        return gs, gt, None, None, None, None


class Attention2Fn(torch.autograd.Function):
    """out = a0 x0 + a1 x1 with (a0, a1) = softmax(w.x0 + b, w.x1 + b): Attention.forward
    (pygda/nn/attention.py:52-55) for two views as one kernel each way."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias):
        x0, x1 = _f32c(x0), _f32c(x1)
        w = _f32c(weight).reshape(-1)
        n, h = x0.shape
        if x1.shape != x0.shape or w.numel() != h:
            raise ValueError("attention: the two views and the weight must agree in width")
        out = torch.empty_like(x0)
        a0 = torch.empty(n, dtype=torch.float32, device=x0.device)
        gda.attention2_fwd(_p(x0), _p(x1), n, h, _p(w), _p(bias), _p(out), _p(a0), _stream())
        ctx.save_for_backward(x0, x1, w, a0)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, go):
        x0, x1, w, a0 = ctx.saved_tensors
        go = _f32c(go)
        n, h = x0.shape
        g0 = torch.empty_like(x0) if ctx.needs_input_grad[0] else None
        g1 = torch.empty_like(x1) if ctx.needs_input_grad[1] else None
        dw = torch.empty(h, dtype=torch.float32, device=x0.device) if ctx.needs_input_grad[2] else None
        gda.attention2_bwd(_p(x0), _p(x1), n, h, _p(w), _p(a0), _p(go), _p(g0), _p(g1), _p(dw), _stream())
        db = None
        if ctx.has_bias and ctx.needs_input_grad[3]:      # the scores share the bias: softmax is shift invariant
            db = torch.empty(1, dtype=torch.float32, device=x0.device)
            gda.fill_f32(_p(db), 1, 0.0, _stream())
        return g0, g1, (dw.view(1, h) if dw is not None else None), db


class SegmentMeanFn(torch.autograd.Function):
    """global_mean_pool over a sorted batch vector given as ptr [G+1]."""

    @staticmethod
    def forward(ctx, x, ptr):
        x = _f32c(x)
        g = ptr.numel() - 1
        out = torch.empty(g, x.shape[1], dtype=torch.float32, device=x.device)
        gda.segment_mean_fwd(_p(x), x.stride(0), _p(ptr), g, x.shape[1], _p(out), _stream())
        ctx.save_for_backward(ptr)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, gout):
        (ptr,) = ctx.saved_tensors
        gout = _f32c(gout)
        gx = torch.empty(ctx.rows, gout.shape[1], dtype=torch.float32, device=gout.device)
        gda.segment_mean_bwd(_p(gout), _p(ptr), ptr.numel() - 1, gout.shape[1], _p(gx), gx.stride(0), _stream())
        return gx, None


def global_mean_pool(x, batch, size=None):
    """PyG ``global_mean_pool`` (a2gnn_base.py:141): ``batch`` sorted graph ids."""
    if batch is None:
        ptr = torch.tensor([0, x.shape[0]], dtype=torch.int64, device=x.device)
        return SegmentMeanFn.apply(x, ptr)
    # the segment pointers are a function of ``batch`` alone: computed once per batch tensor (AdaGCN's step pools the
    # same mini-batch 22 times, adagcn.py:169-198 -- a device->host sync, a bincount and a scan each time otherwise)
    cached = getattr(batch, "_gda_pool_ptr", None)
    if cached is not None and cached[0] == (batch._version, size):
        return SegmentMeanFn.apply(x, cached[1])
    g = int(batch.max().item()) + 1 if size is None else int(size)
    counts = torch.bincount(batch, minlength=g)
    ptr = torch.zeros(g + 1, dtype=torch.int64, device=x.device)
    ptr[1:] = torch.cumsum(counts, 0)
    try:
        batch._gda_pool_ptr = ((batch._version, size), ptr)
    except Exception:                                  # noqa: BLE001 -- exotic tensor subclasses: just do not cache
        pass
    return SegmentMeanFn.apply(x, ptr)


def graph_conv(x, weight, bias, graph, k, w_in_out=False):
    if x.dtype == torch.bfloat16:
        return GraphConvBf16Fn.apply(x, weight, bias, graph, int(k), bool(w_in_out))
    return GraphConvFn.apply(x, weight, bias, graph, int(k), bool(w_in_out))


def linear(x, weight, bias=None):
    return LinearFn.apply(x, weight, bias)


class MatmulFn(torch.autograd.Function):
    """a @ b (or a @ b^T) on libgda, differentiable in both operands: the small dense products of AdaGCN's
    closed-form WGAN-GP penalty (W1 W1^T and u (W1 W1^T), pygda_b200/models/adagcn.py)."""

    @staticmethod
    def forward(ctx, a, b, trans_b):
        a, b = _f32c(a), _f32c(b)
        ctx.save_for_backward(a, b)
        ctx.trans_b = trans_b
        return mm(a, b, trans_b=trans_b)[0]

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _f32c(g)
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = mm(g, b, trans_b=not ctx.trans_b)[0]                    # g b^T  |  g b
        if ctx.needs_input_grad[1]:
            gb = mm(g, a, trans_a=True)[0] if ctx.trans_b else mm(a, g, trans_a=True)[0]   # g^T a  |  a^T g
        return ga, gb, None


def matmul(a, b, trans_b=False):
    return MatmulFn.apply(a, b, bool(trans_b))


def softmax_cross_entropy(logits, labels):
    return SoftmaxCEFn.apply(logits, labels, 0)


def domain_cross_entropy(logits, split):
    """CE against the implicit domain labels [0]*split + [1]*(rows-split) (a2gnn.py:200-204)."""
    return SoftmaxCEFn.apply(logits, None, int(split))


def softmax_entropy(logits):
    return SoftmaxEntropyFn.apply(logits)
