"""A2GNNBase -- drop-in for pygda/nn/a2gnn_base.py:7-203 (same ctor, methods and
state-dict names: ``convs.{i}.lin.weight``, ``convs.{i}.bias``, ``cls.*``,
``domain_discriminator.*``)."""
import os

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .prop_gcn_conv import PropGCNConv


class _Linear(nn.Linear):
    """nn.Linear whose matmul runs on libgda (heads at a2gnn_base.py:67,70)."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class A2GNNBase(nn.Module):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, adv=False, dropout=0.1,
                 act=F.relu, mode='node', **kwargs):
        super().__init__()
        self.in_dim, self.hid_dim, self.num_classes = in_dim, hid_dim, num_classes
        self.num_layers, self.adv, self.dropout = num_layers, adv, dropout
        self.act, self.mode = act, mode
        self.convs = nn.ModuleList()
        self.convs.append(PropGCNConv(in_dim, hid_dim))
        for _ in range(num_layers - 1):
            self.convs.append(PropGCNConv(hid_dim, hid_dim))
        if mode == 'node':
            self.cls = PropGCNConv(hid_dim, num_classes)
        else:
            self.cls = _Linear(hid_dim, num_classes)
        if adv:
            self.domain_discriminator = _Linear(hid_dim, 2)

    def forward(self, data, prop_nums, first_layer=None):
        """a2gnn_base.py:72-104.  ``first_layer`` (optional) is a precomputed output of
        ``self.first_conv`` for the same (data, prop_nums): layer 1 is deterministic in
        (x, W1, b1, A_hat, k) -- dropout only acts after it (:136-138) -- so the
        estimator shares it between the two bottleneck evaluations per domain
        (SURVEY.md Appendix B.2).  Results are identical to recomputing it."""
        if self.mode == 'node':
            x, edge_index, batch = data.x, data.edge_index, None
        else:
            x, edge_index, batch = data.x, data.edge_index, data.batch
        x = self.feat_bottleneck(x, edge_index, batch, prop_nums=prop_nums, first_layer=first_layer)
        return self.feat_classifier(x, edge_index, batch, prop_nums=1)

    def first_conv(self, x, edge_index, prop_nums):
        return self.convs[0](x, edge_index, prop_nums)

    def feat_bottleneck(self, x, edge_index, batch, prop_nums=30, first_layer=None):
        for i, conv in enumerate(self.convs):                                # :135-138
            x = first_layer if (i == 0 and first_layer is not None) else conv(x, edge_index, prop_nums)
            x = ops.act_dropout(x, self.act, self.dropout, self.training)
        if self.mode == 'graph':                                             # :140-141
            x = ops.global_mean_pool(x, batch)
        return x

    def feat_bottleneck_pair(self, x, edge_index, batch, prop_nums=30, first_layer=None):
        """The TWO ``feat_bottleneck`` evaluations the reference makes per domain and step
        (models/a2gnn.py:181 & :192 for the source, :193 & :211 for the target) in one pass:
        layer 1 once (its inputs are identical, dropout acts after it), every later layer on the
        stacked pair with independent dropout masks.  Returns ``(features_1, features_2)`` with the
        values two separate calls would give.  Without dropout the two evaluations coincide and
        one tensor is returned twice."""
        p = float(self.dropout) if self.training else 0.0
        if p == 0.0:
            f = self.feat_bottleneck(x, edge_index, batch, prop_nums=prop_nums, first_layer=first_layer)
            return f, f
        if self.mode != 'node' or not (self.act is F.relu or self.act is torch.relu) or \
                os.environ.get("GDA_NO_PAIR") == "1":             # A/B switch for profiles/
            return (self.feat_bottleneck(x, edge_index, batch, prop_nums=prop_nums, first_layer=first_layer),
                    self.feat_bottleneck(x, edge_index, batch, prop_nums=prop_nums, first_layer=first_layer))
        z1 = first_layer if first_layer is not None else self.convs[0](x, edge_index, prop_nums)
        a, b = ops.act_dropout_pair(z1, p)
        for conv in list(self.convs)[1:]:
            a, b = conv.forward_pair(a, b, edge_index, prop_nums, dropout_p=p)
        return a, b

    def feat_classifier(self, x, edge_index, batch, prop_nums=1):            # :145-176
        if self.mode == 'node':
            return self.cls(x, edge_index, prop_nums)
        return self.cls(x)

    def domain_classifier(self, x, alpha):                                   # :178-203
        return self.domain_discriminator(ops.GradReverse.apply(x, alpha))
