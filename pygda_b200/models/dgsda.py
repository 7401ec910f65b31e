"""DGSDA -- drop-in for pygda/models/dgsda.py:18-436 (ctor :73-118, forward_model :144-196,
entropy_minimization_loss :198-227, fit :229-349, predict :367-436) on the B200 path.

Objective: CE(source) + alpha * L1(temp_s, temp_t) + beta * MMD(relu(lin1 x_s), relu(lin1 x_t)) + gamma * class-
frequency-weighted target entropy.  The propagations (4 x BernProp per step) are ``gda_spmm_f32`` launches, the
MMD / CE / linear layers the kernels of the A2GNN path; the two scalar-ish terms -- L1 over K+1 temperatures and the
weighted entropy over [N, C] logits -- stay tiny torch expressions."""
import torch
import torch.nn.functional as F

from . import BaseGDA
from .. import ops
from ..nn.dgsda_base import DGSDABase
from ..optim import Adam
from ..utils import MMD
from ._common import TwoDomainLoop


class DGSDA(TwoDomainLoop, BaseGDA):
    def __init__(self, in_dim, hid_dim, num_classes, mode='node', num_layers=2, dropout=0., act=F.relu, K=8,
                 alpha=0.05, beta=0.5, gamma=0.05, weight_decay=0., lr=4e-3, epoch=200, device='cuda:0', batch_size=0,
                 num_neigh=-1, verbose=2, **kwargs):
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes, num_layers=num_layers,
                         dropout=dropout, act=act, weight_decay=weight_decay, lr=lr, epoch=epoch, device=device,
                         batch_size=batch_size, num_neigh=num_neigh, verbose=verbose, **kwargs)
        assert num_layers == 2, 'unsupport number of layers'                             # :110
        assert mode == 'node', 'unsupport mode'                                          # :111
        self.K = K
        self.mode = mode
        self.alpha = alpha
        self.beta = beta
        self.gamma = gamma

    def init_model(self, **kwargs):
        return DGSDABase(features=self.in_dim, hidden=self.hid_dim, classes=self.num_classes, dprate=self.dropout,
                         K=self.K, **kwargs).to(self.device)

    def forward_model(self, source_data, target_data, mmd_indices=None):
        net = self.dgsda
        source_logits = net(source_data)                                                  # :178
        train_loss = ops.softmax_cross_entropy(source_logits, source_data.y)              # :179
        theta_loss = F.l1_loss(net.prop1.temp, net.prop2.temp)                            # :182-184
        source_feature = ops.act_dropout(net.lin1(source_data.x), F.relu, 0.0, False)     # :187-188
        target_feature = ops.act_dropout(net.lin1(target_data.x), F.relu, 0.0, False)
        mmd_loss = MMD(source_feature, target_feature, indices=mmd_indices)               # :189
        target_outputs = net(target_data, False)                                          # :192
        entropy_loss = self.entropy_minimization_loss(target_outputs)                     # :193
        loss = train_loss + theta_loss * self.alpha + mmd_loss * self.beta + entropy_loss * self.gamma
        return loss, source_logits

    @staticmethod
    def entropy_minimization_loss(output):                                                # :198-227
        probs = F.softmax(output, dim=1)
        log_probs = F.log_softmax(output, dim=1)
        a = torch.sum(probs, dim=0)
        return -torch.sum(probs * log_probs / (a / torch.sum(a)), dim=1).mean()

    def train_step(self, source_data, target_data, optimizer, mmd_indices=None):
        self.dgsda.train()
        source_data = source_data.to(self.device)
        target_data = target_data.to(self.device)
        loss, source_logits = self.forward_model(source_data, target_data, mmd_indices=mmd_indices)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss, source_logits, source_data

    def fit(self, source_data, target_data):
        self._build_loaders(source_data, target_data)
        self.dgsda = self.init_model(**self.kwargs)
        optimizer = Adam(self.dgsda.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer = optimizer
        self._fit_loop(lambda epoch, s, t: self.train_step(s, t, optimizer))

    def process_graph(self, data):
        pass

    def predict(self, data, source=False):
        self.dgsda.eval()
        return self._predict_loop(lambda d: self.dgsda(d, source), source)
