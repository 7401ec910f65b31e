"""GNNBase -- drop-in for the gcn branch of pygda/nn/gnn_base.py:8-205 (stock ``GCNConv`` stack;
``cls`` is a ``GCNConv`` in node mode, a Linear in graph mode; ``forward`` applies log_softmax,
:137).  ``gat`` builds ``GATConv(heads=1, concat=False)`` (:81-87) on the edge-softmax kernels of gat.cu;
``sage`` / ``gin`` are not on the path named by BASELINE.json."""
import torch.nn.functional as F
from torch import nn

from .. import ops
from .layers import Linear
from .gat_conv import GATConv
from .prop_gcn_conv import GCNConv


class GNNBase(nn.Module):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, dropout=0.1, act=F.relu, gnn='gcn',
                 mode='node', **kwargs):
        super().__init__()
        assert gnn in ('gcn', 'sage', 'gat', 'gin'), 'Invalid gnn backbone'
        if gnn not in ('gcn', 'gat'):
            raise NotImplementedError(f"gnn='{gnn}' is not on the accelerated path (gcn and gat are)")
        self.in_dim, self.hid_dim, self.num_classes = in_dim, hid_dim, num_classes
        self.num_layers, self.dropout, self.gnn, self.act, self.mode = num_layers, dropout, gnn, act, mode
        conv = GCNConv if gnn == 'gcn' else GATConv                 # GATConv(heads=1, concat=False), :81-87
        self.convs = nn.ModuleList()
        self.convs.append(conv(in_dim, hid_dim))
        for _ in range(num_layers - 1):
            self.convs.append(conv(hid_dim, hid_dim))
        self.cls = conv(hid_dim, num_classes) if mode == 'node' else Linear(hid_dim, num_classes)

    def forward(self, x, edge_index, edge_weight=None, batch=None):
        x = self.feat_bottleneck(x, edge_index, edge_weight)
        if self.mode == 'graph':
            x = ops.global_mean_pool(x, batch)
        x = self.feat_classifier(x, edge_index, edge_weight)
        return F.log_softmax(x, dim=1)        # [N, C] with C <= a few classes: tiny, left to torch

    def feat_bottleneck(self, x, edge_index, edge_weight=None):
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index, edge_weight)
            if i < len(self.convs) - 1:
                x = ops.act_dropout(x, self.act, self.dropout, self.training)
        return x

    def feat_classifier(self, x, edge_index, edge_weight=None):
        if self.mode == 'node':
            return self.cls(x, edge_index, edge_weight)
        return self.cls(x)
