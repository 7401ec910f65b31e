"""Device graph handles (``gda_graph_t``) and the per-edge_index cache.

The reference recomputes ``gcn_norm`` on every conv call
(pygda/nn/prop_gcn_conv.py:182-192, ``cached=False``): (2L+1) x 2 rebuilds per
training step of a result that is a pure function of ``edge_index``.  Here the
normalised CSR/CSC is built once per distinct ``edge_index`` (and flag set) and
looked up afterwards -- same values, no per-step work.
"""
import ctypes as C
from collections import OrderedDict

import torch

from ._lib import gda, load

SELF_LOOPS, IMPROVED, NORM_SYM_COL, NORM_SYM_ROW, SKIP_EMPTY_ROWS = 1, 2, 4, 8, 16


def _stream():
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Graph:
    """Owns one ``gda_graph_t``; destroyed with the Python object."""

    def __init__(self, edge_index, num_nodes, edge_weight=None, flags=SELF_LOOPS | NORM_SYM_COL):
        load()
        if not edge_index.is_cuda:
            raise ValueError("pygda_b200 graphs live on the GPU: edge_index must be a CUDA tensor")
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be int64 [2, E]")
        ei = edge_index.contiguous()
        w = None
        if edge_weight is not None:
            w = edge_weight.to(torch.float32).contiguous()
            if w.numel() != ei.size(1):
                raise ValueError("edge_weight must have one entry per edge")
        self.device = ei.device
        self.num_nodes = int(num_nodes)
        self.flags = int(flags)
        self._h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            gda.graph_create(_ptr(ei), ei.size(1), self.num_nodes, _ptr(w), self.flags, _stream(),
                             C.byref(self._h))
        n, nnz, nl, nlt = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        gda.graph_info(self._h, C.byref(n), C.byref(nnz), C.byref(nl), C.byref(nlt))
        self.nnz, self.num_long_rows, self.num_long_rows_t = nnz.value, nl.value, nlt.value
        self._ws = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                load().gda_graph_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p(0)

    @property
    def handle(self):
        return self._h

    def workspace(self, transpose, width):
        """Scratch for the split long rows; cached per (orientation, width)."""
        key = (int(bool(transpose)), int(width))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = load().gda_spmm_workspace_bytes(self._h, key[0], key[1])
            ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # ---- exports used by the parity tests -------------------------------------------------
    def coo(self):
        """(edge_index int64 [2,nnz], weight fp32 [nnz]) in the reference's order."""
        ei = torch.empty(2, self.nnz, dtype=torch.int64, device=self.device)
        w = torch.empty(self.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            gda.graph_export_coo(self._h, _ptr(ei), _ptr(w), _stream())
        return ei, w

    def csr(self, transpose=False):
        rp = torch.empty(self.num_nodes + 1, dtype=torch.int32, device=self.device)
        ci = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
        v = torch.empty(self.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            gda.graph_export_csr(self._h, int(bool(transpose)), _ptr(rp), _ptr(ci), _ptr(v), _stream())
        return rp, ci, v


class GraphCache:
    """LRU of Graph handles keyed by the identity of ``edge_index``.

    The key is ``edge_index._gda_key`` when present (set by ``Data.to`` so a
    graph copied from the same host tensor every step maps to one handle), else
    (data_ptr, version, shape).  The cache keeps the keyed tensor alive, so a
    data_ptr can never be recycled while its entry exists."""

    def __init__(self, capacity=32):
        self.capacity = capacity
        self._d = OrderedDict()

    @staticmethod
    def key_of(edge_index):
        k = getattr(edge_index, "_gda_key", None)
        if k is not None:
            return k
        return ("dev", edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape),
                str(edge_index.device))

    def get(self, edge_index, num_nodes, edge_weight=None, flags=SELF_LOOPS | NORM_SYM_COL):
        wkey = None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version)
        key = (self.key_of(edge_index), int(num_nodes), wkey, int(flags), str(edge_index.device))
        hit = self._d.get(key)
        if hit is not None:
            self._d.move_to_end(key)
            return hit[0]
        g = Graph(edge_index, num_nodes, edge_weight, flags)
        keep = getattr(edge_index, "_gda_keepalive", edge_index)
        self._d[key] = (g, keep, edge_weight)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return g

    def clear(self):
        self._d.clear()


_default_cache = GraphCache()


def graph_for(edge_index, num_nodes, edge_weight=None, flags=SELF_LOOPS | NORM_SYM_COL):
    return _default_cache.get(edge_index, num_nodes, edge_weight, flags)


def clear_graph_cache():
    _default_cache.clear()
