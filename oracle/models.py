"""Oracle restatement of the pygda.models estimators on the hot path (CPU).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only what the parity tests and
the CPU baseline need: ``forward_model`` and a one-step ``train_step`` that
follows the reference's inner loop line by line.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import mmd as M
from . import nn as ONN
from . import pyg_ops as P


class A2GNN:
    """pygda/models/a2gnn.py:66-411 (ctor :66-108, forward_model :146-213,
    inner loop of fit :300-319, predict :356-411)."""

    def __init__(self, in_dim, hid_dim, num_classes, mode="node", num_layers=3, dropout=0.,
                 act=F.relu, s_pnums=0, t_pnums=30, adv=False, weight=5, weight_decay=0.,
                 lr=4e-3, epoch=200, device="cpu", **kwargs):
        self.in_dim, self.hid_dim, self.num_classes = in_dim, hid_dim, num_classes
        self.mode, self.num_layers, self.dropout, self.act = mode, num_layers, dropout, act
        self.s_pnums, self.t_pnums, self.adv, self.weight = s_pnums, t_pnums, adv, weight
        self.weight_decay, self.lr, self.epoch, self.device = weight_decay, lr, epoch, device
        self.a2gnn = ONN.A2GNNBase(in_dim, hid_dim, num_classes, num_layers=num_layers, adv=adv,
                                   dropout=dropout, act=act, mode=mode).to(device)
        self.optimizer = torch.optim.Adam(self.a2gnn.parameters(), lr=lr,
                                          weight_decay=weight_decay)
        self.mmd_sqdist = M.pairwise_sqdist_broadcast
        self.mmd_indices = None     # tests may inject (source_sample, target_sample)

    def forward_model(self, source_data, target_data, alpha):
        net = self.a2gnn
        source_logits = net(source_data, self.s_pnums)                                    # :181
        loss = F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y)             # :182
        sb = None if self.mode == "node" else source_data.batch
        tb = None if self.mode == "node" else target_data.batch
        source_features = net.feat_bottleneck(source_data.x, source_data.edge_index, sb,
                                              self.s_pnums)                              # :192
        target_features = net.feat_bottleneck(target_data.x, target_data.edge_index, tb,
                                              self.t_pnums)                              # :193
        if self.adv:                                                                      # :196-205
            sd = net.domain_classifier(source_features, alpha)
            td = net.domain_classifier(target_features, alpha)
            domain_label = torch.tensor([0] * source_data.x.shape[0] + [1] * target_data.x.shape[0])
            loss = loss + self.weight * F.cross_entropy(torch.cat([sd, td], 0), domain_label)
        else:                                                                             # :207-209
            loss = loss + M.MMD(source_features, target_features, indices=self.mmd_indices,
                                sqdist=self.mmd_sqdist) * self.weight
        target_logits = net(target_data, self.t_pnums)                                    # :211
        return loss, source_logits, target_logits

    @staticmethod
    def alpha_at(epoch, total):
        p = float(epoch) / total                                                          # :305
        return 2. / (1. + np.exp(-10. * p)) - 1                                           # :306

    def train_step(self, source_data, target_data, epoch=0):
        """One iteration of the loop body at pygda/models/a2gnn.py:308-319."""
        self.a2gnn.train()
        loss, s_logits, t_logits = self.forward_model(source_data, target_data,
                                                      self.alpha_at(epoch, self.epoch))
        val = loss.item()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return val, s_logits, t_logits

    def predict(self, data, source=False):
        self.a2gnn.eval()
        with torch.no_grad():
            return self.a2gnn(data, self.s_pnums if source else self.t_pnums), data.y
