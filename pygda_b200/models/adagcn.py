"""AdaGCN -- drop-in for pygda/models/adagcn.py:17-454 (ctor :67-109, forward_model :138-198,
fit :200-319, predict :321-385, gradient_penalty :387-454).

Per step the reference runs 10 critic iterations, each with two encoder forwards whose graph is
kept and back-propagated into the encoder although only the critic's optimiser steps
(:169-183) -- the encoder gradients produced there are cleared by ``optimizer.zero_grad()``
(:302) before anyone reads them.  Here the encoder forwards of the critic loop run without a
tape (same values, same critic updates, 10 encoder backward passes less); the final encoder
forward + backward is unchanged.

The critic (Linear-ReLU-Dropout-Linear-Sigmoid on [3N, hid] rows in node mode, [<= 3B, hid] in graph mode) runs on
libgda by default (``analytic_critic = True``): both of its Linear layers through ``ops.linear``, its Adam through
``pygda_b200.optim.Adam``, and the WGAN-GP gradient penalty (:387-454) in CLOSED FORM instead of by
``autograd.grad(create_graph=True)``.  For D(x) = sigmoid(w2 . (relu(W1 x + b1) * m) + b2) with dropout mask m,

    dD/dx = D (1 - D) * (u W1),   u = w2 * 1[W1 x + b1 > 0] * m          (a [rows, adv_dim] matrix)
    |dD/dx|^2 = (D (1 - D))^2 * rowsum((u W1 W1^T) * u)

so no [rows, hid] gradient tensor is ever formed and first-order autograd over [rows, adv_dim]
arrays yields the same parameter gradients (relu' and the mask are piecewise constant, exactly as in
torch's double backward); the two dense products W1 W1^T and u (W1 W1^T) go through ``ops.matmul``.  What is left to
torch are elementwise expressions over [rows, adv_dim] / [rows] arrays.  ``analytic_critic = False`` keeps the
reference's literal form (``torch.nn`` critic, double backward) for comparison (tests/test_zz3_gpu_critic.py: same
values and gradients as torch's double backward).  Checked against the double-backward form (tests/test_zz3_gpu_critic.py)."""
import torch
import torch.nn.functional as F
from torch import nn

from . import BaseGDA
from .. import ops
from ..nn.adagcn_base import AdaGCNBase
from ..optim import Adam
from ._common import TwoDomainLoop


class AdaGCN(TwoDomainLoop, BaseGDA):
    def __init__(self, in_dim, hid_dim, num_classes, mode='node', num_layers=3, dropout=0., act=F.relu,
                 gnn_type='gcn', adv_dim=40, gp_weight=5, domain_weight=1, weight_decay=0., lr=4e-3, epoch=100,
                 device='cuda:0', batch_size=0, num_neigh=-1, verbose=2, **kwargs):
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes, num_layers=num_layers,
                         dropout=dropout, act=act, weight_decay=weight_decay, lr=lr, epoch=epoch, device=device,
                         batch_size=batch_size, num_neigh=num_neigh, verbose=verbose, **kwargs)
        self.gnn_type = gnn_type
        self.adv_dim = adv_dim
        self.gp_weight = gp_weight
        self.domain_weight = domain_weight
        self.mode = mode
        self.analytic_critic = True       # closed-form penalty, critic on libgda; see the module docstring

    def init_model(self, **kwargs):
        return AdaGCNBase(in_dim=self.in_dim, hid_dim=self.hid_dim, num_classes=self.num_classes,
                          num_layers=self.num_layers, dropout=self.dropout, act=self.act, gnn_type=self.gnn_type,
                          mode=self.mode, **kwargs).to(self.device)

    def init_critic(self):
        """The critic and its optimiser, created inside ``fit`` by the reference (:264-276)."""
        self.discriminator = nn.Sequential(nn.Linear(self.hid_dim, self.adv_dim), nn.ReLU(), nn.Dropout(0.1),
                                           nn.Linear(self.adv_dim, 1), nn.Sigmoid()).to(self.device)
        self.c_optimizer = Adam(self.discriminator.parameters(), lr=self.lr, weight_decay=self.weight_decay)

    # ---- critic evaluation -------------------------------------------------------------------------------
    def _critic_hidden(self, x):
        """(z1, u-mask, h) of the critic's first block: z1 = W1 x + b1 on libgda; m = dropout keep mask / (1 - p)
        in train mode; h = relu(z1) * m."""
        lin1, drop = self.discriminator[0], self.discriminator[2]
        z1 = ops.linear(x, lin1.weight, lin1.bias)
        pos = (z1 > 0).to(z1.dtype)
        if self.discriminator.training and drop.p > 0:
            pos = pos * ((torch.rand_like(z1) >= drop.p).to(z1.dtype) / (1.0 - drop.p))
        return z1, pos, z1 * pos

    def _critic(self, x):
        """``self.discriminator(x)`` -- through libgda for the [rows, hid] product when ``analytic_critic``."""
        if not self.analytic_critic:
            return self.discriminator(x)
        _, _, h = self._critic_hidden(x)
        lin2 = self.discriminator[3]
        return torch.sigmoid(ops.linear(h, lin2.weight, lin2.bias))

    def _gradient_penalty_closed_form(self, inputs):
        lin1, lin2 = self.discriminator[0], self.discriminator[3]
        _, pos, h = self._critic_hidden(inputs.detach())
        score = torch.sigmoid(ops.linear(h, lin2.weight, lin2.bias)).reshape(-1)   # D(x)            [rows]
        u = pos * lin2.weight.reshape(1, -1)                               # w2 * relu' * dropout      [rows, adv]
        gram = ops.matmul(lin1.weight, lin1.weight, trans_b=True)          # W1 W1^T                   [adv, adv]
        q = (ops.matmul(u, gram) * u).sum(dim=1)                           # |u W1|^2                  [rows]
        live = q > 0                                                       # norm(0) has gradient 0 in torch
        root = torch.where(live, q, torch.ones_like(q)).sqrt() * live.to(q.dtype)
        gradient_norm = (score * (1 - score)).abs() * root
        return torch.mean((gradient_norm - 1) ** 2)

    def forward_model(self, source_data, target_data):
        for _ in range(10):                                                               # :169-183
            with torch.no_grad():                       # see the module docstring
                encoded_source = self.adagcn(source_data)
                encoded_target = self.adagcn(target_data)
            gp_loss = self.gradient_penalty(encoded_source, encoded_target)
            dis_s = torch.mean(self._critic(encoded_source).reshape(-1))
            dis_t = torch.mean(self._critic(encoded_target).reshape(-1))
            dis_loss = - torch.abs(dis_s - dis_t)
            loss = dis_loss + self.gp_weight * gp_loss
            self.c_optimizer.zero_grad()
            loss.backward()
            self.c_optimizer.step()
        encoded_source = self.adagcn(source_data)                                         # :185-186
        encoded_target = self.adagcn(target_data)
        source_logits = self.adagcn.cls_model(encoded_source)
        cls_loss = self.adagcn.loss_func(source_logits, source_data.y)
        dis_s = torch.mean(self._critic(encoded_source).reshape(-1))
        dis_t = torch.mean(self._critic(encoded_target).reshape(-1))
        dis_loss = torch.abs(dis_s - dis_t)
        target_logits = self.adagcn.cls_model(encoded_target)
        loss = cls_loss + dis_loss * self.domain_weight                                   # :196
        return loss, source_logits, target_logits

    def _rand(self, n):
        # CPU generator, like the reference (:420,:429,:434); through pinned memory so that the copy does not stall
        # the host behind everything queued on the GPU
        from ..utils.mmd import to_device_async
        return to_device_async(torch.rand((n, 1)), torch.device(self.device))

    def gradient_penalty(self, encoded_source, encoded_target):
        num_s, num_t = encoded_source.shape[0], encoded_target.shape[0]                   # :387-454
        if num_s < num_t:
            hidden = encoded_target[-num_s:, ]
            hidden_s = torch.cat((encoded_source, encoded_source), dim=0)
            hidden_t = torch.cat((encoded_target[0:num_s, ], hidden), dim=0)
            alpha = self._rand(2 * num_s)
            interpolates = hidden_t + (alpha * (hidden_s - hidden_t))
        elif num_s > num_t:
            hidden = encoded_source[-num_t:, ]
            hidden_s = torch.cat((encoded_source[0:num_t, ], hidden), dim=0)
            hidden_t = torch.cat((encoded_target, encoded_target), dim=0)
            alpha = self._rand(2 * num_t)
            interpolates = hidden_t + (alpha * (hidden_s - hidden_t))
        else:
            alpha = self._rand(num_t)
            interpolates = encoded_target + (alpha * (encoded_source - encoded_target))
        inputs = torch.cat((encoded_source, encoded_target, interpolates), dim=0)
        if self.analytic_critic:
            return self._gradient_penalty_closed_form(inputs)
        if not inputs.requires_grad:                    # tape-free encoder outputs: differentiate w.r.t. a leaf
            inputs = inputs.detach().requires_grad_(True)
        scores = self.discriminator(inputs)
        gradient = torch.autograd.grad(inputs=inputs, outputs=scores,
                                       grad_outputs=torch.ones_like(scores).to(self.device),
                                       create_graph=True, retain_graph=True, only_inputs=True)[0]
        gradient = gradient.view(gradient.shape[0], -1)
        gradient_norm = gradient.norm(2, dim=1)
        return torch.mean((gradient_norm - 1) ** 2)

    def train_step(self, source_data, target_data, optimizer):
        self.adagcn.train()
        source_data = source_data.to(self.device)
        target_data = target_data.to(self.device)
        loss, source_logits, target_logits = self.forward_model(source_data, target_data)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss, source_logits, target_logits, source_data

    def fit(self, source_data, target_data):
        self._build_loaders(source_data, target_data)
        self.adagcn = self.init_model(**self.kwargs)
        optimizer = Adam(self.adagcn.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer = optimizer
        self.init_critic()

        def step(epoch, s, t):
            loss, source_logits, _, s = self.train_step(s, t, optimizer)
            return loss, source_logits, s

        self._fit_loop(step)

    def process_graph(self, data):
        pass

    def predict(self, data, source=False):
        self.adagcn.eval()
        return self._predict_loop(lambda d: self.adagcn.cls_model(self.adagcn(d)), source)
