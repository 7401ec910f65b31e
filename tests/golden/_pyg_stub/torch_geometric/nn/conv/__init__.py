import inspect

import torch

from oracle import pyg_ops as P


class MessagePassing(torch.nn.Module):
    """node_dim=0; aggr 'add' | 'mean'; flow 'source_to_target' (x_j = x[edge_index[0]], reduce at
    edge_index[1]) | 'target_to_source' (x_j = x[edge_index[1]], reduce at edge_index[0])
    (SURVEY Appendix A.2).  'mean' = sum / max(count, 1) as in PyG's MeanAggregation."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=0, **kwargs):
        super().__init__()
        assert aggr in ('add', 'mean') and flow in ('source_to_target', 'target_to_source') and node_dim == 0
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        assert torch.is_tensor(edge_index)
        j, i = (0, 1) if self.flow == 'source_to_target' else (1, 0)
        params = list(inspect.signature(self.message).parameters)
        args = {}
        n = None if size is None else size[i]
        for name in params:
            if name.endswith('_j'):
                src = kwargs[name[:-2]]
                n = src.size(0) if n is None else n
                args[name] = src.index_select(0, edge_index[j])
            elif name.endswith('_i'):
                src = kwargs[name[:-2]]
                n = src.size(0) if n is None else n
                args[name] = src.index_select(0, edge_index[i])
            elif name == 'edge_index':
                args[name] = edge_index
            else:
                args[name] = kwargs.get(name)
        msg = self.message(**args)
        out = P.scatter_add(msg, edge_index[i], 0, n)
        if self.aggr == 'mean':
            cnt = P.scatter_add(torch.ones(edge_index.size(1), dtype=msg.dtype), edge_index[i], 0, n)
            out = out / cnt.clamp(min=1).view(-1, 1)
        extra = {k: kwargs[k] for k in list(inspect.signature(self.update).parameters)[1:] if k in kwargs}
        return self.update(out, **extra)

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
