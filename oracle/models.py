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


class TDSS(A2GNN):
    """pygda/models/tdss.py:93-312 (ctor :168-214, forward_model :241-312, smoothness :314-383,
    compute_laplacian_loss :385-449, TwoHopNeighbor :21-90).  SURVEY.md section 8(f) row 1."""

    def __init__(self, in_dim, hid_dim, num_classes, mode="node", smooth_mode="RW", num_layers=2, dropout=0.,
                 act=F.relu, s_pnums=0, t_pnums=30, k=2, rw_len=4, alpha=0.001, beta=1e-4, weight_decay=0.005,
                 adv=False, lr=0.01, epoch=200, device="cpu", **kwargs):
        assert mode == "node", "TDSS only supports node-level tasks"                      # :205
        assert adv == False, "TDSS does not support adversarial training"  # noqa: E712   # :206
        super().__init__(in_dim, hid_dim, num_classes, mode=mode, num_layers=num_layers, dropout=dropout, act=act,
                         s_pnums=s_pnums, t_pnums=t_pnums, adv=adv, weight=alpha, weight_decay=weight_decay, lr=lr,
                         epoch=epoch, device=device, **kwargs)
        self.smooth_mode, self.k, self.rw_len, self.alpha, self.beta = smooth_mode, k, rw_len, alpha, beta
        self.rw_generator = None

    @staticmethod
    def two_hop(edge_index, num_nodes):
        """TwoHopNeighbor.__call__ (:66-90) for ``edge_attr is None``."""
        value = edge_index.new_ones((edge_index.size(1),), dtype=torch.float)
        index, value = P.spspmm(edge_index, value, edge_index, value, num_nodes, num_nodes, num_nodes, True)   # :73
        value.fill_(0)
        index, value = P.remove_self_loops(index, value)                                  # :75
        edge_index = torch.cat([edge_index, index], dim=1)                                # :77
        out, _ = P.coalesce(edge_index, None, num_nodes)                                  # :79
        return out

    def smoothness(self, edge_index, edge_attr, num_nodes):                               # :314-383
        if self.smooth_mode == "RW":
            row, col = edge_index
            start = torch.arange(num_nodes)
            walk = P.random_walk(row, col, start, self.rw_len, generator=self.rw_generator)    # :370
            adj = torch.zeros((num_nodes, num_nodes), dtype=torch.float)
            adj[walk[start], start.unsqueeze(1)] = 1.0                                    # :372
            return P.dense_to_sparse(adj)                                                 # :373
        if self.k == 1:
            return P.add_remaining_self_loops(edge_index, edge_attr)                      # :376
        assert edge_attr is None, "oracle restates the edge_attr=None branch"
        hop = edge_index
        for _ in range(self.k - 1):                                                       # :381-382
            hop = self.two_hop(hop, num_nodes)
        return P.add_remaining_self_loops(hop, None, num_nodes=num_nodes)                 # :385

    @staticmethod
    def compute_laplacian_loss(features, edge_index):                                     # :385-449
        edge_weight = torch.ones(edge_index.size(1))
        row, col = edge_index
        deg = torch.zeros(features.size(0)).scatter_add_(0, row, edge_weight)
        dinv = deg.pow(-0.5)
        dinv[torch.isinf(dinv)] = 0
        diff = features[row] * dinv[row].view(-1, 1) - features[col] * dinv[col].view(-1, 1)
        return ((diff).pow(2).sum(dim=1) * edge_weight).sum() / 2.

    def forward_model(self, source_data, target_data, alpha):                             # :241-312
        net = self.a2gnn
        source_logits = net(source_data, self.s_pnums)
        loss = F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y)
        source_features = net.feat_bottleneck(source_data.x, source_data.edge_index, None, self.s_pnums)
        target_features = net.feat_bottleneck(target_data.x, target_data.edge_index, None, self.t_pnums)
        mmd_loss = M.MMD(source_features, target_features, indices=self.mmd_indices, sqdist=self.mmd_sqdist)
        loss = loss + self.alpha * mmd_loss                                               # :300
        laplacian_loss = self.compute_laplacian_loss(target_features, target_data.edge_index_smooth)   # :303
        loss = loss + self.beta * laplacian_loss                                          # :304
        target_logits = net(target_data, self.t_pnums)                                    # :306
        return loss, source_logits, target_logits


class UDAGCN:
    """pygda/models/udagcn.py:64-308 with ``ppmi=False`` (forward_model :131-201, loop body
    :277-293).  Note the encoder's dropout layers are always active (oracle/nn.py)."""

    def __init__(self, in_dim, hid_dim, num_classes, mode="node", num_layers=2, adv_dim=40,
                 weight_decay=3e-3, lr=4e-3, epoch=300, act=F.relu, ppmi=False, **kwargs):
        import itertools
        self.mode, self.epoch = mode, epoch
        self.udagcn = ONN.UDAGCNBase(in_dim, hid_dim, num_classes, num_layers=num_layers, act=act,
                                     ppmi=ppmi, adv_dim=adv_dim)
        params = itertools.chain(*[m.parameters() for m in self.udagcn.models])
        self.optimizer = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)

    def forward_model(self, source_data, target_data, alpha, epoch):
        net = self.udagcn
        es = net.encode(source_data, "source")
        et = net.encode(target_data, "target")
        if self.mode == "graph":
            es = P.global_mean_pool(es, source_data.batch)
            et = P.global_mean_pool(et, target_data.batch)
        source_logits = net.cls_model(es)
        cls_loss = net.loss_func(source_logits, source_data.y)
        sd = net.domain_model(ONN.GradReverse.apply(es, alpha))
        td = net.domain_model(ONN.GradReverse.apply(et, alpha))
        loss_grl = net.loss_func(sd, torch.zeros(sd.size(0)).type(torch.LongTensor)) + \
            net.loss_func(td, torch.ones(td.size(0)).type(torch.LongTensor))
        loss = cls_loss + loss_grl
        target_logits = net.cls_model(et)
        probs = torch.clamp(F.softmax(target_logits, dim=-1), min=1e-9, max=1.0)
        loss_entropy = torch.mean(torch.sum(-probs * torch.log(probs), dim=-1))
        loss = loss + loss_entropy * (epoch / self.epoch * 0.01)
        return loss, source_logits, target_logits

    def train_step(self, source_data, target_data, epoch=0):
        for m in self.udagcn.models:
            m.train()
        alpha = min((epoch + 1) / self.epoch, 0.05)
        loss, s_logits, t_logits = self.forward_model(source_data, target_data, alpha, epoch)
        val = loss.item()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return val, s_logits, t_logits


class GRADE:
    """pygda/models/grade.py:62-301 (forward_model :129-197, loop body :271-287)."""

    def __init__(self, in_dim, hid_dim, num_classes, mode="node", num_layers=2, dropout=0.,
                 act=F.relu, disc="JS", weight=0.01, weight_decay=0.01, lr=0.001, epoch=200, **kwargs):
        self.mode, self.disc, self.weight, self.epoch = mode, disc, weight, epoch
        self.grade = ONN.GRADEBase(in_dim, hid_dim, num_classes, num_layers=num_layers,
                                   dropout=dropout, act=act, disc=disc, mode=mode)
        self.optimizer = torch.optim.Adam(self.grade.parameters(), lr=lr, weight_decay=weight_decay)
        self.mmd_indices, self.mmd_sqdist = None, M.pairwise_sqdist_broadcast

    def _num(self, d):
        return d.x.size(0) if self.mode == "node" else len(d)

    def forward_model(self, source_data, target_data, alpha):
        net = self.grade
        source_logits, source_feats = net(source_data)
        target_logits, target_feats = net(target_data)
        loss = F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y)
        domain_loss = 0
        labels = torch.tensor([0] * self._num(source_data) + [1] * self._num(target_data))
        if self.disc == "JS":
            preds = net.discriminator(ONN.GradReverse.apply(torch.cat([source_feats, target_feats], 0), alpha))
            domain_loss = net.criterion(preds, labels)
        elif self.disc == "MMD":
            mind = min(self._num(source_data), self._num(target_data))
            domain_loss = M.MMD(source_feats[:mind], target_feats[:mind], indices=self.mmd_indices,
                                sqdist=self.mmd_sqdist)
        elif self.disc == "C":
            s_l_f = torch.cat([source_feats, 8 * net.one_hot_embedding(source_data.y)], dim=1)
            t_l_f = torch.cat([target_feats, 8 * F.softmax(target_logits, dim=1)], dim=1)
            preds = net.discriminator(ONN.GradReverse.apply(torch.cat([s_l_f, t_l_f], 0), alpha))
            domain_loss = net.criterion(preds, labels)
        return loss + domain_loss * self.weight, source_logits, target_logits

    def train_step(self, source_data, target_data, epoch=0):
        self.grade.train()
        alpha = 2 / (1 + np.exp(-10 * epoch / self.epoch)) - 1
        loss, s_logits, t_logits = self.forward_model(source_data, target_data, alpha)
        val = loss.item()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return val, s_logits, t_logits


class AdaGCN:
    """pygda/models/adagcn.py:67-454: forward_model :138-198 (10 critic iterations with WGAN-GP,
    encoder graph kept as in the reference), critic :264-276, gradient_penalty :387-454."""

    def __init__(self, in_dim, hid_dim, num_classes, mode="node", num_layers=3, adv_dim=40, gp_weight=5,
                 domain_weight=1, weight_decay=0., lr=4e-3, act=F.relu, **kwargs):
        self.mode, self.gp_weight, self.domain_weight = mode, gp_weight, domain_weight
        self.adagcn = ONN.AdaGCNBase(in_dim, hid_dim, num_classes, num_layers=num_layers, act=act, mode=mode)
        self.optimizer = torch.optim.Adam(self.adagcn.parameters(), lr=lr, weight_decay=weight_decay)
        self.discriminator = torch.nn.Sequential(
            torch.nn.Linear(hid_dim, adv_dim), torch.nn.ReLU(), torch.nn.Dropout(0.1),
            torch.nn.Linear(adv_dim, 1), torch.nn.Sigmoid())
        self.c_optimizer = torch.optim.Adam(self.discriminator.parameters(), lr=lr, weight_decay=weight_decay)

    def gradient_penalty(self, es, et):
        num_s, num_t = es.shape[0], et.shape[0]
        if num_s < num_t:
            hs = torch.cat((es, es), 0)
            ht = torch.cat((et[0:num_s, ], et[-num_s:, ]), 0)
            alpha = torch.rand((2 * num_s, 1))
            inter = ht + alpha * (hs - ht)
        elif num_s > num_t:
            hs = torch.cat((es[0:num_t, ], es[-num_t:, ]), 0)
            ht = torch.cat((et, et), 0)
            alpha = torch.rand((2 * num_t, 1))
            inter = ht + alpha * (hs - ht)
        else:
            alpha = torch.rand((num_t, 1))
            inter = et + alpha * (es - et)
        inputs = torch.cat((es, et, inter), 0)
        scores = self.discriminator(inputs)
        grad = torch.autograd.grad(inputs=inputs, outputs=scores, grad_outputs=torch.ones_like(scores),
                                   create_graph=True, retain_graph=True, only_inputs=True)[0]
        return torch.mean((grad.view(grad.shape[0], -1).norm(2, dim=1) - 1) ** 2)

    def forward_model(self, source_data, target_data):
        for _ in range(10):
            es, et = self.adagcn(source_data), self.adagcn(target_data)
            gp = self.gradient_penalty(es, et)
            dis_loss = -torch.abs(torch.mean(self.discriminator(es).reshape(-1)) -
                                  torch.mean(self.discriminator(et).reshape(-1)))
            loss = dis_loss + self.gp_weight * gp
            self.c_optimizer.zero_grad()
            loss.backward()
            self.c_optimizer.step()
        es, et = self.adagcn(source_data), self.adagcn(target_data)
        source_logits = self.adagcn.cls_model(es)
        cls_loss = self.adagcn.loss_func(source_logits, source_data.y)
        dis_loss = torch.abs(torch.mean(self.discriminator(es).reshape(-1)) -
                             torch.mean(self.discriminator(et).reshape(-1)))
        target_logits = self.adagcn.cls_model(et)
        return cls_loss + dis_loss * self.domain_weight, source_logits, target_logits


class GNN:
    """pygda/models/gnn.py:59-268 over the gcn backbone (forward_model :120-150)."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=3, dropout=0., act=F.relu, gnn="gcn", **kwargs):
        self.gnn = ONN.GNNBase(in_dim, hid_dim, num_classes, num_layers=num_layers, dropout=dropout, act=act,
                               gnn=gnn)

    def forward_model(self, source_data, target_data):
        source_logits = self.gnn(source_data.x, source_data.edge_index)
        target_logits = self.gnn(target_data.x, target_data.edge_index)
        return F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y), source_logits, target_logits


class DGSDA:
    """pygda/models/dgsda.py:73-227 (forward_model :144-196, entropy_minimization_loss :198-227, loop body
    :300-330): CE(source) + alpha * L1(temp_s, temp_t) + beta * MMD(relu(lin1 x_s), relu(lin1 x_t)) +
    gamma * weighted target entropy."""

    def __init__(self, in_dim, hid_dim, num_classes, mode="node", num_layers=2, dropout=0., K=8, alpha=0.05,
                 beta=0.5, gamma=0.05, weight_decay=0., lr=4e-3, epoch=200, **kwargs):
        assert num_layers == 2 and mode == "node"
        self.alpha, self.beta, self.gamma, self.epoch = alpha, beta, gamma, epoch
        self.dgsda = ONN.DGSDABase(in_dim, hid_dim, num_classes, dprate=dropout, K=K)
        self.optimizer = torch.optim.Adam(self.dgsda.parameters(), lr=lr, weight_decay=weight_decay)
        self.mmd_indices, self.mmd_sqdist = None, M.pairwise_sqdist_broadcast

    @staticmethod
    def entropy_minimization_loss(output):
        probs = F.softmax(output, dim=1)
        log_probs = F.log_softmax(output, dim=1)
        a = torch.sum(probs, dim=0)
        return -torch.sum(probs * log_probs / (a / torch.sum(a)), dim=1).mean()

    def forward_model(self, source_data, target_data):
        net = self.dgsda
        source_logits = net(source_data)
        loss = F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y)
        loss = loss + F.l1_loss(net.prop1.temp, net.prop2.temp) * self.alpha
        source_feature = F.relu(net.lin1(source_data.x))
        target_feature = F.relu(net.lin1(target_data.x))
        loss = loss + M.MMD(source_feature, target_feature, indices=self.mmd_indices, sqdist=self.mmd_sqdist) * self.beta
        target_outputs = net(target_data, False)
        loss = loss + self.entropy_minimization_loss(target_outputs) * self.gamma
        return loss, source_logits

    def train_step(self, source_data, target_data):
        self.dgsda.train()
        loss, s_logits = self.forward_model(source_data, target_data)
        val = loss.item()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return val, s_logits


class StruRW:
    """pygda/models/strurw.py:21-758 -- ctor :83-139, init_model :141-187, forward_model :189-257,
    forward_model_mixup :259-323, loop body of fit :406-421, cal_reweight :446-487, cal_edge_prob_sep :489-548,
    predict :669-700, shuffle_data / id_node :702-758."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=2, cls_dim=128, cls_layers=2, dropout=0., gnn="GS",
                 pooling="mean", reweight=True, pseudo=True, ew_start=100, ew_freq=20, lamb=0.8, mode="erm",
                 act=F.relu, bn=False, weight_decay=0.0001, lr=0.05, epoch=100, **kwargs):
        assert mode in ["erm", "mixup", "mmd", "adv"], "unsupport training mode"
        self.num_classes, self.mode, self.epoch = num_classes, mode, epoch
        self.reweight, self.pseudo, self.ew_start, self.ew_freq = reweight, pseudo, ew_start, ew_freq
        if mode == "mixup":
            self.gnn = ONN.MixupBase(in_dim, hid_dim, num_classes, num_layers=num_layers, dropout=dropout,
                                     rw_lmda=lamb)
        else:
            self.gnn = ONN.ReweightGNN(in_dim, hid_dim, num_classes, cls_dim, gnn_layers=num_layers,
                                       cls_layers=cls_layers, backbone=gnn, pooling=pooling, dropout=dropout, bn=bn,
                                       rw_lmda=lamb)
        params = list(self.gnn.parameters())
        if mode == "adv":
            self.domain_discriminator = torch.nn.Linear(hid_dim, 2)
            params += list(self.domain_discriminator.parameters())
        self.optimizer = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)
        self.mmd_indices, self.mmd_sqdist = None, M.pairwise_sqdist_broadcast

    # ---- edge re-weighting -----------------------------------------------------------------
    @staticmethod
    def class_edge_counts(edge_index, labels, num_classes):
        """``one_hot.T * to_dense_adj(edge_index) * one_hot`` (:533-535): entry [i, j] = number of edges (with
        multiplicity -- to_dense_adj sums duplicates) whose edge_index[0] end has class i and whose
        edge_index[1] end has class j.  Counted directly instead of through the dense N x N adjacency; the
        values are exact integers either way."""
        code = labels[edge_index[0]] * num_classes + labels[edge_index[1]]
        return torch.bincount(code, minlength=num_classes * num_classes).view(num_classes, num_classes).double()

    def cal_edge_prob_sep(self, src_graph, tgt_graph, tgt_pred):           # :489-548
        C = self.num_classes
        n_src = torch.bincount(src_graph.y, minlength=C).double()
        n_pred = torch.bincount(tgt_pred, minlength=C).double()
        n_tgt = torch.bincount(tgt_graph.y, minlength=C).double()
        src_edge_prob = self.class_edge_counts(src_graph.edge_index, src_graph.y, C) / torch.outer(n_src, n_src)
        tgt_edge_prob = self.class_edge_counts(tgt_graph.edge_index, tgt_pred, C) / (torch.outer(n_pred, n_pred) + 1e-12)
        tgt_true_edge_prob = self.class_edge_counts(tgt_graph.edge_index, tgt_graph.y, C) / torch.outer(n_tgt, n_tgt)
        return src_edge_prob, tgt_edge_prob, tgt_true_edge_prob

    def cal_reweight(self, source_data, target_data, target_pred):         # :446-487
        src_edge_prob, tgt_edge_prob, _ = self.cal_edge_prob_sep(source_data, target_data, target_pred)
        reweight_matrix = torch.div(tgt_edge_prob, src_edge_prob)
        reweight_matrix[torch.isinf(reweight_matrix)] = 1
        reweight_matrix[torch.isnan(reweight_matrix)] = 1
        # :479-485 -- edge e gets reweight_matrix[i][j] with i = class of edge_index[1][e], j = class of
        # edge_index[0][e] (both source labels: the indices of a source edge never reach the target_pred part
        # of ``label_pred``); float64 -> float32 on assignment
        y = source_data.y
        ei = source_data.edge_index
        source_data.edge_weight = reweight_matrix[y[ei[1]], y[ei[0]]].float()

    def _maybe_reweight(self, source_data, target_data, target_pred, epoch):   # :220-226 / :283-289
        if self.reweight and (epoch + 1) >= self.ew_start:
            if self.pseudo:
                if (epoch + 1) % self.ew_freq == 0:
                    self.cal_reweight(source_data, target_data, target_pred)
            elif epoch == self.ew_start - 1:
                self.cal_reweight(source_data, target_data, target_pred)

    # ---- objectives ------------------------------------------------------------------------
    def forward_model(self, source_data, target_data, alpha, epoch):       # :189-257
        target_feat, target_logits = self.gnn.forward(target_data, target_data.x)
        target_pred = torch.max(F.softmax(target_logits, dim=1), dim=1)[1]
        self._maybe_reweight(source_data, target_data, target_pred, epoch)
        source_feat, source_logits = self.gnn.forward(source_data, source_data.x)
        loss = F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y)
        if self.mode == "adv":
            sd = self.domain_discriminator(ONN.GradReverse.apply(source_feat, alpha))
            td = self.domain_discriminator(ONN.GradReverse.apply(target_feat, alpha))
            domain_label = torch.tensor([0] * source_data.x.shape[0] + [1] * target_data.x.shape[0])
            loss = loss + F.cross_entropy(torch.cat([sd, td], 0), domain_label)
        elif self.mode == "mmd":
            loss = loss + M.MMD(source_feat, target_feat, indices=self.mmd_indices, sqdist=self.mmd_sqdist)
        return loss, source_logits, target_logits

    @staticmethod
    def shuffle_edges(edge_index, num_nodes):
        """shuffle_data / id_node (:702-758): a node permutation from ``np.random.shuffle`` and the edge list
        re-labelled through its inverse."""
        id_new_value_old = np.arange(num_nodes)
        np.random.shuffle(id_new_value_old)
        perm = torch.from_numpy(id_new_value_old)
        id_old_value_new = torch.zeros(num_nodes, dtype=torch.long)
        id_old_value_new[perm] = torch.arange(num_nodes, dtype=torch.long)
        return torch.stack([id_old_value_new[edge_index[0]], id_old_value_new[edge_index[1]]], dim=0), id_new_value_old

    def forward_model_mixup(self, source_data, target_data, epoch):        # :259-323
        n_t = target_data.x.shape[0]
        target_feat = self.gnn.feat_bottleneck(target_data.x, target_data.edge_index, target_data.edge_index, 1,
                                               np.arange(n_t), target_data.edge_weight)
        target_logits = self.gnn.feat_classifier(target_feat)
        target_pred = torch.max(F.softmax(target_logits, dim=1), dim=1)[1]
        self._maybe_reweight(source_data, target_data, target_pred, epoch)
        lam = np.random.beta(4.0, 4.0)
        edge_index_b, id_new_value_old = self.shuffle_edges(source_data.edge_index, source_data.x.shape[0])
        source_feat = self.gnn.feat_bottleneck(source_data.x, source_data.edge_index, edge_index_b, lam,
                                               id_new_value_old, source_data.edge_weight)
        source_logits = self.gnn.feat_classifier(source_feat)
        # :321 -- the loss uses the UNMIXED labels only (kept as in the reference)
        loss = F.nll_loss(F.log_softmax(source_logits, dim=1), source_data.y)
        return loss, source_logits, target_logits

    def train_step(self, source_data, target_data, epoch=0):               # :406-421
        self.gnn.train()
        if self.mode == "mixup":
            loss, s_logits, t_logits = self.forward_model_mixup(source_data, target_data, epoch)
        else:
            alpha = A2GNN.alpha_at(epoch, self.epoch)
            loss, s_logits, t_logits = self.forward_model(source_data, target_data, alpha, epoch)
        val = loss.item()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return val, s_logits, t_logits

    def predict(self, data):                                               # :669-700
        self.gnn.eval()
        with torch.no_grad():
            if self.mode == "mixup":
                data.edge_weight = torch.ones(data.edge_index.shape[1])
                logits = self.gnn(data.x, data.edge_index, data.edge_index, 1, np.arange(data.x.shape[0]),
                                  data.edge_weight)
            else:
                _, logits = self.gnn(data, data.x)
        return logits, data.y
