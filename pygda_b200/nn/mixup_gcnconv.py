"""MixUpGCNConv -- drop-in for pygda/nn/mixup_gcnconv.py:91-247.

out = sum_{e: col_e = c} w_e ((1 - lmda) + lmda rw_e) (x W^T)[row_e] + x_cen W_cen^T + b, with w = gcn_norm of unit
weights WITHOUT self loops (:199-206): the GCN path's fused GEMM + aggregation node on a re-weighted CSR
(``message_graph``) plus one Linear."""
import torch
from torch import nn

from .. import ops
from .prop_gcn_conv import GlorotLinear
from .reweight_gnn import message_graph


class MixUpGCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=False,
                 normalize=True, bias=True, **kwargs):
        super().__init__()
        if improved or add_self_loops or not normalize or kwargs.get('aggr', 'add') != 'add':
            raise NotImplementedError("MixUpGCNConv is only ever built with its defaults (mixup_base.py:62-65)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached, self.add_self_loops, self.normalize = improved, cached, add_self_loops, normalize
        self.lin = GlorotLinear(in_channels, out_channels)
        self.lin_cen = GlorotLinear(in_channels, out_channels)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)

    def reset_parameters(self):                                               # :145-149 (lin_cen is NOT reset there)
        self.lin.reset_parameters()
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x, x_cen, edge_index, edge_weight=None, lmda=1):        # :151-215
        if edge_weight is None:
            raise TypeError("MixUpGCNConv needs edge weights (the reference's message calls edge_rw.view, :243)")
        g = message_graph(edge_index, edge_weight, lmda, x.size(0), True, 'add', to_source=False)
        return ops.add(ops.graph_conv(x, self.lin.weight, None, g, 1), ops.linear(x_cen, self.lin_cen.weight, self.bias))
