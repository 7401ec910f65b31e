"""AdaGCNBase -- drop-in for pygda/nn/adagcn_base.py:9-181: ``GNN`` encoder (stock ``GCNConv`` stack,
act + Dropout between layers, ``global_mean_pool`` in graph mode) + ``cls_model`` Linear.
``gnn_type='ppmi'`` swaps in ``PPMIConv`` layers (adagcn_base.py:53-57; PPMI graph built on the GPU)."""
import torch.nn.functional as F
from torch import nn

from .. import ops
from .layers import Linear
from .ppmi_conv import PPMIConv
from .prop_gcn_conv import GCNConv


class GNN(nn.Module):
    def __init__(self, in_dim, hid_dim, gnn_type='gcn', num_layers=3, act=F.relu, dropout=0.1, **kwargs):
        super().__init__()
        self.gnn_type, self.act, self.num_layers = gnn_type, act, num_layers
        # adagcn_base.py:48-57: 'gcn' -> stock GCNConv, anything else -> PPMIConv (whose graph is cached under
        # "default_cache" forever -- in graph mode the first batch's PPMI graph is re-used, as in the reference)
        conv_cls = GCNConv if gnn_type == 'gcn' else PPMIConv
        self.conv_layers = nn.ModuleList()
        self.conv_layers.append(conv_cls(in_dim, hid_dim))
        for _ in range(1, num_layers):
            self.conv_layers.append(conv_cls(hid_dim, hid_dim))
        self.dropout = nn.Dropout(dropout)          # holds p and the train/eval flag; applied fused below

    def forward(self, x, edge_index, batch, mode='node'):
        for i, conv_layer in enumerate(self.conv_layers):
            x = conv_layer(x, edge_index)
            if i < len(self.conv_layers) - 1:
                x = ops.act_dropout(x, self.act, self.dropout.p, self.dropout.training)
        if mode == 'graph':
            x = ops.global_mean_pool(x, batch)
        return x


class AdaGCNBase(nn.Module):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=3, dropout=0.1, act=F.relu, gnn_type='gcn',
                 mode='node', **kwargs):
        super().__init__()
        # the reference does not forward `dropout` to the encoder (adagcn_base.py:84): it keeps 0.1
        self.encoder = GNN(in_dim=in_dim, hid_dim=hid_dim, gnn_type=gnn_type, act=act, num_layers=num_layers)
        self.cls_model = nn.Sequential(Linear(hid_dim, num_classes))
        self.mode = mode
        self.loss_func = ops.softmax_cross_entropy

    def forward(self, data):
        if self.mode == 'node':
            x, edge_index, batch = data.x, data.edge_index, None
        else:
            x, edge_index, batch = data.x, data.edge_index, data.batch
        return self.encoder(x, edge_index, batch, mode=self.mode)
