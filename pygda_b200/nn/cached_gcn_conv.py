"""CachedGCNConv -- drop-in for pygda/nn/cached_gcn_conv.py:7-177.

``weight`` is [in, out] (``x @ W``, :49,130), weight/bias may be shared ``Parameter``s
passed in by the caller (:35-61), the normalised graph is cached per ``cache_name`` forever
(:132-136 -- including the reference's quirk that a later batch with different edges re-uses
the first batch's graph), degree is accumulated at the SOURCE index (:98-103)."""
import math

import torch
from torch import nn

from .. import ops
from ..graph import IMPROVED, NORM_SYM_ROW, SELF_LOOPS, Graph


class CachedGCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, weight=None, bias=None, improved=False,
                 use_bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.improved = in_channels, out_channels, improved
        self.cache_dict = {}
        if weight is None:
            self.weight = nn.Parameter(torch.empty(in_channels, out_channels, dtype=torch.float32))
            a = math.sqrt(6.0 / (in_channels + out_channels))           # PyG glorot
            with torch.no_grad():
                self.weight.uniform_(-a, a)
        else:
            self.weight = weight
        if bias is None:
            if use_bias:
                self.bias = nn.Parameter(torch.zeros(out_channels, dtype=torch.float32))
            else:
                self.register_parameter('bias', None)
        else:
            self.bias = bias

    @staticmethod
    def norm(edge_index, num_nodes, edge_weight=None, improved=False, dtype=None):
        """Same contract as the reference's static ``norm`` (:63-103)."""
        flags = SELF_LOOPS | NORM_SYM_ROW | (IMPROVED if improved else 0)
        return Graph(edge_index, num_nodes, edge_weight, flags).coo()

    def _graph(self, edge_index, num_nodes, cache_name, edge_weight=None):
        if cache_name not in self.cache_dict:
            flags = SELF_LOOPS | NORM_SYM_ROW | (IMPROVED if self.improved else 0)
            self.cache_dict[cache_name] = Graph(edge_index, num_nodes, edge_weight, flags)
        return self.cache_dict[cache_name]

    def forward(self, x, edge_index, cache_name="default_cache", edge_weight=None):
        graph = self._graph(edge_index, x.size(0), cache_name, edge_weight)
        return ops.graph_conv(x, self.weight, self.bias, graph, 1, w_in_out=True)

    def propagate_product(self, xw, edge_index, cache_name="default_cache", edge_weight=None):
        """``forward`` given ``xw = x @ self.weight`` computed by the caller: aggregation + bias only.  Lets two
        layers that share ``weight`` AND input (UDAGCN's adjacency and PPMI views, udagcn_base.py:168-169) share
        the product -- same values as two ``forward`` calls."""
        graph = self._graph(edge_index, xw.size(0), cache_name, edge_weight)
        out = ops.PropagateFn.apply(xw, graph, 1)
        return out if self.bias is None else ops.BiasAddFn.apply(out, self.bias)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)
