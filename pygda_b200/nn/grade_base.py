"""GRADEBase -- drop-in for pygda/nn/grade_base.py:8-202: stock ``GCNConv`` stack, per-layer
features collected, ``cls`` Linear, features = cat([h_1..h_L, logits]) of width L*hid + C."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .layers import Linear
from .prop_gcn_conv import GCNConv


class GRADEBase(nn.Module):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, dropout=0.1, act=F.relu, disc='JS',
                 mode='node', **kwargs):
        super().__init__()
        self.in_dim, self.hid_dim, self.num_classes = in_dim, hid_dim, num_classes
        self.num_layers, self.dropout, self.act, self.mode = num_layers, dropout, act, mode
        self.convs = nn.ModuleList()
        self.convs.append(GCNConv(in_dim, hid_dim))
        for _ in range(num_layers - 1):
            self.convs.append(GCNConv(hid_dim, hid_dim))
        self.cls = Linear(hid_dim, num_classes)
        width = hid_dim * num_layers + (num_classes if disc == "JS" else num_classes * 2)
        self.discriminator = nn.Sequential(Linear(width, 2))
        self.criterion = ops.softmax_cross_entropy

    def forward(self, data):
        if self.mode == 'node':
            x, edge_index, batch = data.x, data.edge_index, None
        else:
            x, edge_index, batch = data.x, data.edge_index, data.batch
        x, feat_list = self.feat_bottleneck(x, edge_index, batch)
        x = self.feat_classifier(x)
        feat_list.append(x)
        return x, torch.cat(feat_list, dim=1)

    def feat_bottleneck(self, x, edge_index, batch):
        feat_list = []
        for conv in self.convs:
            x = conv(x, edge_index)
            x = ops.act_dropout(x, self.act, self.dropout, self.training)
            feat_list.append(x if self.mode == 'node' else ops.global_mean_pool(x, batch))
        if self.mode == 'graph':
            x = ops.global_mean_pool(x, batch)
        return x, feat_list

    def feat_classifier(self, x):
        return self.cls(x)

    def one_hot_embedding(self, labels):
        return torch.eye(self.num_classes, device=labels.device)[labels]
