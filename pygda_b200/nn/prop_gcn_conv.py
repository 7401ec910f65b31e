"""PropGCNConv -- drop-in for pygda/nn/prop_gcn_conv.py:84-264.

Same constructor, parameter names (``lin.weight`` [out,in] glorot, ``bias`` zeros)
and ``forward(x, edge_index, prop_nums=1, edge_weight=None)``; the body is one
fused tape node  y = A_hat^k (x W^T) + b  on libgda kernels.  ``gcn_norm``
(prop_gcn_conv.py:24-81) is evaluated once per distinct ``edge_index`` instead of
on every call (the reference's ``cached=False`` recomputation returns the same
values each time).
"""
import math

import torch
from torch import nn

from .. import ops
from ..graph import IMPROVED, NORM_SYM_COL, SELF_LOOPS, graph_for


class GlorotLinear(nn.Module):
    """PyG ``Linear(in, out, bias=False, weight_initializer='glorot')``
    (prop_gcn_conv.py:136-137): weight [out, in], y = x W^T."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        with torch.no_grad():
            self.weight.uniform_(-a, a)

    def forward(self, x):
        return ops.linear(x, self.weight, None)


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             dtype=None):
    """Same contract as the tensor branch of prop_gcn_conv.py:64-81: returns the
    self-looped ``edge_index`` and normalised weights, in the reference's order."""
    if num_nodes is None:
        num_nodes = int(edge_index.max().item()) + 1 if edge_index.numel() else 0
    flags = NORM_SYM_COL | (SELF_LOOPS if add_self_loops else 0) | (IMPROVED if improved else 0)
    return graph_for(edge_index, num_nodes, edge_weight, flags).coo()


class PropGCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self._cached_graph = None
        self.lin = GlorotLinear(in_channels, out_channels)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def reset_parameters(self):
        self.lin.reset_parameters()
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()
        self._cached_graph = None

    def _graph(self, x, edge_index, edge_weight):
        if self.cached and self._cached_graph is not None:     # prop_gcn_conv.py:182-192
            return self._cached_graph
        flags = 0
        if self.normalize:
            flags = NORM_SYM_COL | (SELF_LOOPS if self.add_self_loops else 0) | \
                (IMPROVED if self.improved else 0)
        part = getattr(edge_index, "_gda_partition", None)
        if part is not None:               # row-partitioned multi-GPU graph (pygda_b200/dist.py)
            g = part.graph(edge_index, flags, edge_weight)
        else:
            g = graph_for(edge_index, x.size(0), edge_weight, flags)
        if self.cached:
            self._cached_graph = g
        return g

    def forward(self, x, edge_index, prop_nums=1, edge_weight=None):
        k = int(prop_nums)
        graph = self._graph(x, edge_index, edge_weight) if k > 0 else None
        return ops.graph_conv(x, self.lin.weight, self.bias, graph, k)

    def forward_pair(self, xa, xb, edge_index, prop_nums=1, dropout_p=0.0, edge_weight=None):
        """``dropout(relu(self(x, edge_index, prop_nums)))`` for two inputs at once, with independent masks:
        what two consecutive ``feat_bottleneck`` evaluations of a layer compute (a2gnn_base.py:135-138),
        as one stacked GEMM and one batched aggregation per step (ops.PairGraphConvActFn)."""
        k = int(prop_nums)
        graph = self._graph(xa, edge_index, edge_weight) if k > 0 else None
        if graph is not None and hasattr(graph, "spmm_k") and not graph.supports_nb(self.out_channels, 2):
            # partitioned graph outside halo mode: one matrix per exchange
            import torch.nn.functional as F
            return tuple(ops.act_dropout(ops.graph_conv(x, self.lin.weight, self.bias, graph, k), F.relu,
                                         dropout_p, True) for x in (xa, xb))
        return ops.graph_conv_act_pair(xa, xb, self.lin.weight, self.bias, graph, k, dropout_p)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


class GCNConv(PropGCNConv):
    """Stock PyG ``GCNConv`` (call sites pygda/nn/grade_base.py:58-61, adagcn_base.py:49-52,
    gnn_base.py:65-71) = one propagation step."""

    def forward(self, x, edge_index, edge_weight=None):
        return super().forward(x, edge_index, 1, edge_weight)
