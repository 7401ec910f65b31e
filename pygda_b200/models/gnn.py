"""GNN -- drop-in for pygda/models/gnn.py:18-268 (source-only baseline over ``GNNBase``).  Like the
reference, ``fit`` does not move the batches to the device (:258-262 has no ``.to``) -- the data must
already live there -- and ``predict(data)`` really uses ``data`` (:264-268)."""
import torch
import torch.nn.functional as F

from . import BaseGDA
from .. import ops
from ..data import NeighborLoader
from ..metrics import micro_f1_from_logits
from ..nn.gnn_base import GNNBase
from ..optim import Adam
from ..utils import logger
import time


class GNN(BaseGDA):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=2, dropout=0., gnn='gcn', act=F.relu,
                 weight_decay=0.0001, lr=0.05, epoch=100, device='cuda:0', batch_size=0, num_neigh=-1, verbose=2,
                 **kwargs):                                                               # gnn.py:59-75
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes, num_layers=num_layers,
                         dropout=dropout, act=act, weight_decay=weight_decay, lr=lr, epoch=epoch, device=device,
                         batch_size=batch_size, num_neigh=num_neigh, verbose=verbose, **kwargs)
        self.gnn_type = gnn
        self.gnn = gnn            # the reference keeps the backbone NAME here until fit() replaces it by the module (:92)

    def init_model(self, **kwargs):
        return GNNBase(in_dim=self.in_dim, hid_dim=self.hid_dim, num_classes=self.num_classes,
                       num_layers=self.num_layers, dropout=self.dropout, act=self.act, gnn=self.gnn_type,
                       **kwargs).to(self.device)

    def forward_model(self, source_data, target_data):
        source_logits = self.gnn(source_data.x, source_data.edge_index)                  # :143-144
        target_logits = self.gnn(target_data.x, target_data.edge_index)
        # the model already returns log_softmax (gnn_base.py:137); the loss applies it again (:146)
        loss = ops.softmax_cross_entropy(source_logits, source_data.y)
        return loss, source_logits, target_logits

    def fit(self, source_data, target_data):
        bs_s = source_data.x.shape[0] if self.batch_size == 0 else self.batch_size
        bs_t = target_data.x.shape[0] if self.batch_size == 0 else self.batch_size
        source_loader = NeighborLoader(source_data, self.num_neigh, batch_size=bs_s)
        target_loader = NeighborLoader(target_data, self.num_neigh, batch_size=bs_t)
        self.gnn = self.init_model(**self.kwargs)
        optimizer = Adam(self.gnn.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        start_time = time.time()
        for epoch in range(self.epoch):
            epoch_loss = 0
            logits_all = labels_all = None
            for idx, (s, t) in enumerate(zip(source_loader, target_loader)):
                self.gnn.train()
                loss, _, _ = self.forward_model(s, t)
                epoch_loss += loss.item()
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                if self.verbose > 1:                    # the reference re-predicts every step only to print F1
                    lg, lb = self.predict(s)
                    logits_all = lg if idx == 0 else torch.cat((logits_all, lg))
                    labels_all = lb if idx == 0 else torch.cat((labels_all, lb))
            f1 = micro_f1_from_logits(labels_all, logits_all) if self.verbose > 1 else None
            logger(epoch=epoch, loss=epoch_loss, source_train_acc=f1, time=time.time() - start_time,
                   verbose=self.verbose, train=True)

    def process_graph(self, data):
        pass

    def predict(self, data):
        self.gnn.eval()
        with torch.no_grad():
            logits = self.gnn(data.x, data.edge_index)
        return logits, data.y
