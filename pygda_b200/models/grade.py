"""GRADE -- drop-in for pygda/models/grade.py:15-372 (ctor :62-100, forward_model :129-197,
fit :199-301, predict :303-372)."""
import numpy as np
import torch
import torch.nn.functional as F

from . import BaseGDA
from .. import ops
from ..nn import GRADEBase
from ..optim import Adam
from ..utils import MMD
from ._common import TwoDomainLoop


class GRADE(TwoDomainLoop, BaseGDA):
    def __init__(self, in_dim, hid_dim, num_classes, mode='node', num_layers=2, dropout=0., act=F.relu,
                 disc='JS', weight=0.01, weight_decay=0.01, lr=0.001, epoch=200, device='cuda:0', batch_size=0,
                 num_neigh=-1, verbose=2, **kwargs):
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes, num_layers=num_layers,
                         dropout=dropout, act=act, weight_decay=weight_decay, lr=lr, epoch=epoch,
                         device=device, batch_size=batch_size, num_neigh=num_neigh, verbose=verbose, **kwargs)
        self.disc = disc
        self.weight = weight
        self.mode = mode

    def init_model(self, **kwargs):
        return GRADEBase(in_dim=self.in_dim, hid_dim=self.hid_dim, num_classes=self.num_classes,
                         num_layers=self.num_layers, dropout=self.dropout, act=self.act, disc=self.disc,
                         mode=self.mode, **kwargs).to(self.device)

    def _num(self, data):
        return data.x.size(0) if self.mode == 'node' else len(data)

    def forward_model(self, source_data, target_data, alpha, mmd_indices=None):
        net = self.grade
        source_logits, source_feats = net(source_data)                                    # :163
        target_logits, target_feats = net(target_data)                                    # :164
        train_loss = ops.softmax_cross_entropy(source_logits, source_data.y)              # :165
        n_s = self._num(source_data)
        if self.disc == 'JS':                                                             # :169-176
            feats = torch.cat([source_feats, target_feats], dim=0)
            domain_preds = net.discriminator(ops.GradReverse.apply(feats, alpha))
            domain_loss = ops.domain_cross_entropy(domain_preds, n_s)
        elif self.disc == 'MMD':                                                          # :177-182
            mind = min(n_s, self._num(target_data))
            domain_loss = MMD(source_feats[:mind], target_feats[:mind], indices=mmd_indices)
        elif self.disc == 'C':                                                            # :183-193
            ratio = 8
            s_l_f = torch.cat([source_feats, ratio * net.one_hot_embedding(source_data.y)], dim=1)
            t_l_f = torch.cat([target_feats, ratio * F.softmax(target_logits, dim=1)], dim=1)
            domain_preds = net.discriminator(ops.GradReverse.apply(torch.cat([s_l_f, t_l_f], dim=0), alpha))
            domain_loss = ops.domain_cross_entropy(domain_preds, n_s)
        else:
            return train_loss, source_logits, target_logits       # reference: domain_loss stays 0 (:167)
        loss = ops.combine([(train_loss, 1.0), (domain_loss, float(self.weight))])        # :195
        return loss, source_logits, target_logits

    def train_step(self, source_data, target_data, alpha, optimizer, mmd_indices=None):
        self.grade.train()
        source_data = source_data.to(self.device)
        target_data = target_data.to(self.device)
        loss, source_logits, target_logits = self.forward_model(source_data, target_data, alpha,
                                                                mmd_indices=mmd_indices)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss, source_logits, target_logits, source_data

    def fit(self, source_data, target_data):
        self._build_loaders(source_data, target_data)
        self.grade = self.init_model(**self.kwargs)
        optimizer = Adam(self.grade.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer = optimizer

        def step(epoch, s, t):
            alpha = 2 / (1 + np.exp(- 10 * epoch / self.epoch)) - 1                        # :271
            loss, source_logits, _, s = self.train_step(s, t, alpha, optimizer)
            return loss, source_logits, s

        self._fit_loop(step)

    def process_graph(self, data):
        pass

    def predict(self, data, source=False):
        self.grade.eval()
        return self._predict_loop(lambda d: self.grade(d)[0], source)
