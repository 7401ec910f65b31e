"""GATConv(in, out, heads=1, concat=False) -- the stock PyG layer pygda/nn/gnn_base.py:81-87 builds
(SURVEY.md Appendix A.4): ``lin_src`` [out, in] glorot (shared for source and target), ``att_src`` /
``att_dst`` [1, 1, out] glorot, ``bias`` zeros, negative_slope 0.2, fresh self loops, softmax over the
in-edges of every target.  Parameter names follow PyG 2.4 (``lin_src.weight`` ...)."""
import math

import torch
from torch import nn

from .. import ops
from ..graph import SELF_LOOPS, graph_for
from .prop_gcn_conv import GlorotLinear


class _GATAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, a_src, a_dst, graph, slope):
        import ctypes as C
        from .._lib import gda
        h, a_src, a_dst = ops._f32c(h), ops._f32c(a_src).reshape(-1), ops._f32c(a_dst).reshape(-1)
        n, c = h.shape
        out = torch.empty_like(h)
        alpha = torch.empty(graph.nnz, dtype=torch.float32, device=h.device)
        gda.gat_fwd(graph.handle, ops._p(h), c, ops._p(a_src), ops._p(a_dst), float(slope), ops._p(out),
                    ops._p(alpha), ops._stream())
        ctx.save_for_backward(h, a_src, a_dst, alpha)
        ctx.graph, ctx.slope = graph, slope
        return out

    @staticmethod
    def backward(ctx, gout):
        from .._lib import gda
        h, a_src, a_dst, alpha = ctx.saved_tensors
        gout = ops._f32c(gout)
        n, c = h.shape
        dh, das, dad = torch.empty_like(h), torch.empty_like(a_src), torch.empty_like(a_dst)
        scratch = torch.empty_like(alpha)
        gda.gat_bwd(ctx.graph.handle, ops._p(h), c, ops._p(a_src), ops._p(a_dst), float(ctx.slope), ops._p(alpha),
                    ops._p(gout), ops._p(dh), ops._p(das), ops._p(dad), ops._p(scratch), ops._stream())
        return dh, das.view(n, 1), dad.view(n, 1), None, None


class GATConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=False, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True, **kwargs):
        super().__init__()
        if heads != 1 or concat or dropout != 0.0 or not add_self_loops:
            raise NotImplementedError("only the configuration the reference uses: heads=1, concat=False")
        self.in_channels, self.out_channels, self.negative_slope = in_channels, out_channels, negative_slope
        self.lin_src = GlorotLinear(in_channels, out_channels)
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        a = math.sqrt(6.0 / (1 + out_channels))           # PyG glorot on [1, 1, out]: size(-2) + size(-1)
        with torch.no_grad():
            self.att_src.uniform_(-a, a)
            self.att_dst.uniform_(-a, a)

    def forward(self, x, edge_index, edge_weight=None):
        h = self.lin_src(x)
        c = self.out_channels
        a_s = ops.linear(h, self.att_src.view(1, c))       # [N, 1] = <h_i, att_src>
        a_d = ops.linear(h, self.att_dst.view(1, c))
        graph = graph_for(edge_index, x.size(0), None, SELF_LOOPS)       # remove + add self loops, no norm
        out = _GATAggregate.apply(h, a_s, a_d, graph, self.negative_slope)
        if self.bias is not None:
            out = ops.BiasAddFn.apply(out, self.bias)
        return out
