"""TDSS's smoothing graph and Laplacian regulariser on libgda.

``khop_edge_index`` / ``random_walk_edge_index`` replace ``TDSS.smoothness``
(pygda/models/tdss.py:314-383); ``laplacian_loss`` replaces ``TDSS.compute_laplacian_loss``
(:385-449).  See pygda_b200/csrc/smooth.cu for the kernels.
"""
import ctypes as C

import torch

from . import ops
from ._lib import gda, load
from .graph import graph_for


def _export(handle, device):
    try:
        e = int(load().gda_edges_size(handle))
        out = torch.empty(2, e, dtype=torch.int64, device=device)
        gda.edges_export(handle, ops._p(out), ops._stream())
        torch.cuda.current_stream(device).synchronize()
        return out
    finally:
        load().gda_edges_destroy(handle)


def _device_edges(edge_index):
    if not edge_index.is_cuda:
        raise ValueError("pygda_b200 builds the smoothing graph on the GPU: edge_index must be a CUDA tensor")
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must be int64 [2, E]")
    return edge_index.contiguous()


def khop_edge_index(edge_index, num_nodes, k=2):
    """``smooth_mode='K-hop'``: (k-1) x TwoHopNeighbor, then add_remaining_self_loops (tdss.py:374-385)."""
    ei = _device_edges(edge_index)
    h = C.c_void_p(0)
    with torch.cuda.device(ei.device):
        gda.khop_create(ops._p(ei), ei.size(1), int(num_nodes), int(k), ops._stream(), C.byref(h))
        return _export(h, ei.device)


def random_walk_edge_index(edge_index, num_nodes, walk_length=4, seed=None):
    """``smooth_mode='RW'`` (tdss.py:367-373) without the dense N x N matrix.  ``seed`` defaults to a
    draw from the CPU generator, so ``torch.manual_seed`` makes the walks reproducible."""
    ei = _device_edges(edge_index)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    h = C.c_void_p(0)
    with torch.cuda.device(ei.device):
        gda.rw_create(ops._p(ei), ei.size(1), int(num_nodes), int(walk_length), int(seed) & 0xFFFFFFFFFFFFFFFF,
                      ops._stream(), C.byref(h))
        return _export(h, ei.device)


class _SmoothGraph:
    """Unnormalised graph of ``edge_index_smooth`` (unit weights, duplicates and loops kept) + degrees."""

    def __init__(self, edge_index, num_nodes):
        self.graph = graph_for(edge_index, num_nodes, None, 0)
        dev = edge_index.device
        self.out_deg = torch.empty(num_nodes, dtype=torch.float32, device=dev)
        self.in_deg = torch.empty(num_nodes, dtype=torch.float32, device=dev)
        gda.graph_degrees(self.graph.handle, ops._p(self.out_deg), ops._p(self.in_deg), ops._stream())


def _smooth_graph(edge_index, num_nodes):
    g = graph_for(edge_index, num_nodes, None, 0)
    sg = getattr(g, "_smooth", None)
    if sg is None:
        sg = g._smooth = _SmoothGraph(edge_index, num_nodes)
    return sg


class LaplacianFn(torch.autograd.Function):
    """loss = 1/2 sum_e |g[row] - g[col]|^2, g = D^-1/2 f; the gradient is produced by the forward."""

    @staticmethod
    def forward(ctx, features, sg):
        f = ops._f32c(features)
        n, h = f.shape
        g = torch.empty_like(f)
        gda.row_scale_rsqrt_f32(ops._p(f), f.stride(0), ops._p(sg.out_deg), ops._p(g), n, h, ops._stream())
        u_in = ops.spmm(sg.graph, g)
        u_out = ops.spmm(sg.graph, g, transpose=True)
        loss = torch.empty((), dtype=torch.float32, device=f.device)
        df = torch.empty_like(f)
        ws = ops._workspace(load().gda_laplacian_workspace_bytes(), f.device)
        gda.laplacian_finish_f32(ops._p(g), ops._p(u_in), ops._p(u_out), ops._p(sg.out_deg), ops._p(sg.in_deg), n, h,
                                 ops._p(loss), ops._p(df), ops._p(ws), ws.numel(), ops._stream())
        ctx.save_for_backward(df)
        return loss

    @staticmethod
    def backward(ctx, grad):
        (df,) = ctx.saved_tensors
        return ops._scale_by_scalar(df, grad), None


def laplacian_loss(features, edge_index_smooth):
    """``TDSS.compute_laplacian_loss(features, edge_index)`` (tdss.py:385-449)."""
    if not edge_index_smooth.is_cuda:
        edge_index_smooth = edge_index_smooth.to(features.device)
    return LaplacianFn.apply(features, _smooth_graph(edge_index_smooth, features.size(0)))
