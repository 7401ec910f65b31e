import inspect

import torch

from oracle import pyg_ops as P


class MessagePassing(torch.nn.Module):
    """aggr='add', flow='source_to_target', node_dim=0."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=0, **kwargs):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target' and node_dim == 0
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        assert torch.is_tensor(edge_index)
        params = list(inspect.signature(self.message).parameters)
        args = {}
        n = None
        for name in params:
            if name.endswith('_j'):
                src = kwargs[name[:-2]]
                n = src.size(0)
                args[name] = src.index_select(0, edge_index[0])
            elif name.endswith('_i'):
                src = kwargs[name[:-2]]
                n = src.size(0)
                args[name] = src.index_select(0, edge_index[1])
            else:
                args[name] = kwargs.get(name)
        msg = self.message(**args)
        out = P.scatter_add(msg, edge_index[1], 0, n)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


class GCNConv(MessagePassing):
    """Stock GCNConv."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kwargs):
        super().__init__(aggr='add')
        from ..dense.linear import Linear
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.add_self_loops, self.normalize = improved, add_self_loops, normalize
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer='glorot')
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = P.gcn_norm_by_col(
                edge_index, edge_weight, x.size(0), self.improved, self.add_self_loops, x.dtype)
        x = self.lin(x)
        out = self.propagate(edge_index, x=x, edge_weight=edge_weight)
        return out if self.bias is None else out + self.bias

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j
