"""UDAGCN -- drop-in for pygda/models/udagcn.py:17-387 (ctor :64-102, forward_model :131-201,
fit :203-308, predict :310-387)."""
import itertools

import torch.nn.functional as F

from . import BaseGDA
from .. import ops
from ..nn import UDAGCNBase
from ..optim import Adam
from ._common import TwoDomainLoop


class UDAGCN(TwoDomainLoop, BaseGDA):
    def __init__(self, in_dim, hid_dim, num_classes, mode='node', num_layers=2, dropout=0., act=F.relu,
                 ppmi=True, adv_dim=40, weight_decay=3e-3, lr=4e-3, epoch=300, device='cuda:0', batch_size=0,
                 num_neigh=-1, verbose=2, **kwargs):
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes, num_layers=num_layers,
                         dropout=dropout, act=act, weight_decay=weight_decay, lr=lr, epoch=epoch,
                         device=device, batch_size=batch_size, num_neigh=num_neigh, verbose=verbose, **kwargs)
        self.ppmi = ppmi
        self.adv_dim = adv_dim
        self.mode = mode
        # pygda_b200 extension (BASELINE.json config 3, "bf16"): feature_dtype=torch.bfloat16 keeps the input
        # features and the encoder activations in bf16 (parameters, accumulation and losses stay fp32)
        self.feature_dtype = self.kwargs.pop('feature_dtype', None)

    def init_model(self, **kwargs):
        kwargs.pop('feature_dtype', None)
        return UDAGCNBase(in_dim=self.in_dim, hid_dim=self.hid_dim, num_classes=self.num_classes,
                          num_layers=self.num_layers, dropout=self.dropout, act=self.act, ppmi=self.ppmi,
                          adv_dim=self.adv_dim, feature_dtype=self.feature_dtype, **kwargs).to(self.device)

    def forward_model(self, source_data, target_data, alpha, epoch):
        net = self.udagcn
        encoded_source = net.encode(source_data, "source")                               # :163
        encoded_target = net.encode(target_data, "target")                               # :164
        if self.mode == 'graph':                                                          # :168-170
            encoded_source = ops.global_mean_pool(encoded_source, source_data.batch)
            encoded_target = ops.global_mean_pool(encoded_target, target_data.batch)
        source_logits = net.cls_model(encoded_source)                                     # :171
        cls_loss = net.loss_func(source_logits, source_data.y)                            # :172-175
        source_domain_preds = net.domain_model(ops.GradReverse.apply(encoded_source, alpha))
        target_domain_preds = net.domain_model(ops.GradReverse.apply(encoded_target, alpha))
        # all-zero / all-one label vectors built on the host every step (:180-187) folded into
        # the CE kernel as a row threshold
        source_domain_cls_loss = ops.domain_cross_entropy(source_domain_preds, source_domain_preds.size(0))
        target_domain_cls_loss = ops.domain_cross_entropy(target_domain_preds, 0)
        target_logits = net.cls_model(encoded_target)                                     # :193
        loss_entropy = ops.softmax_entropy(target_logits)                                 # :194-197
        loss = ops.combine([(cls_loss, 1.0), (source_domain_cls_loss, 1.0), (target_domain_cls_loss, 1.0),
                            (loss_entropy, epoch / self.epoch * 0.01)])                   # :189-199
        return loss, source_logits, target_logits

    def _set_train(self, flag):
        for model in self.udagcn.models:
            model.train(flag)

    def train_step(self, source_data, target_data, alpha, epoch, optimizer):
        self._set_train(True)
        source_data = source_data.to(self.device)
        target_data = target_data.to(self.device)
        loss, source_logits, target_logits = self.forward_model(source_data, target_data, alpha, epoch)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss, source_logits, target_logits, source_data

    def fit(self, source_data, target_data):
        self._build_loaders(source_data, target_data)
        self.udagcn = self.init_model(**self.kwargs)
        params = itertools.chain(*[model.parameters() for model in self.udagcn.models])    # :262
        optimizer = Adam(params, lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer = optimizer

        def step(epoch, s, t):
            alpha = min((epoch + 1) / self.epoch, 0.05)                                    # :277
            loss, source_logits, _, s = self.train_step(s, t, alpha, epoch, optimizer)
            return loss, source_logits, s

        self._fit_loop(step)

    def process_graph(self, data):
        pass

    def predict(self, data, source=False):
        self._set_train(False)
        name = 'source' if source else 'target'

        def logits(d):
            enc = self.udagcn.encode(d, name)
            if self.mode == 'graph':
                enc = ops.global_mean_pool(enc, d.batch)
            return self.udagcn.cls_model(enc)

        return self._predict_loop(logits, source)
