"""A2GNN -- drop-in for pygda/models/a2gnn.py:18-411 on the B200 path.

Same constructor (:66-108), ``init_model`` (:110-144), ``forward_model(source_data,
target_data, alpha) -> (loss, source_logits, target_logits)`` (:146-213), ``fit``
(:215-336) and ``predict(data, source=False)`` (:356-411, which -- like the
reference -- ignores ``data`` and re-iterates the loaders stored by ``fit``).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import BaseGDA
from .. import ops
from ..nn import A2GNNBase
from ..optim import Adam
from ..utils import MMD
from ._common import TwoDomainLoop


class A2GNN(TwoDomainLoop, BaseGDA):
    def __init__(self, in_dim, hid_dim, num_classes, mode='node', num_layers=3, dropout=0.,
                 act=F.relu, s_pnums=0, t_pnums=30, adv=False, weight=5, weight_decay=0.,
                 lr=4e-3, epoch=200, device='cuda:0', batch_size=0, num_neigh=-1, verbose=2,
                 **kwargs):
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes,
                         num_layers=num_layers, dropout=dropout, act=act,
                         weight_decay=weight_decay, lr=lr, epoch=epoch, device=device,
                         batch_size=batch_size, num_neigh=num_neigh, verbose=verbose, **kwargs)
        self.s_pnums = s_pnums
        self.t_pnums = t_pnums
        self.adv = adv
        self.weight = weight
        self.mode = mode
        self.overlap_streams = True
        self._side = None
        # full-batch node-level fit() replays the loop body from a CUDA graph, with the per-epoch host->device copy
        # of a host-resident graph double-buffered behind it (models/graphed.py); False: issue every kernel eagerly
        self.cuda_graph = True

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def init_model(self, **kwargs):
        return A2GNNBase(in_dim=self.in_dim, hid_dim=self.hid_dim, num_classes=self.num_classes,
                         num_layers=self.num_layers, adv=self.adv, dropout=self.dropout,
                         act=self.act, mode=self.mode, **kwargs).to(self.device)

    def forward_model(self, source_data, target_data, alpha, mmd_indices=None):
        net = self.a2gnn
        if self.mode == 'node':
            source_batch = target_batch = None
        else:
            source_batch, target_batch = source_data.batch, target_data.batch

        # The source branch (s_pnums is usually 0: dense GEMMs, DRAM-bound) and the target branch
        # (k propagation steps, L2-latency-bound) are independent until the domain loss, so they
        # are issued on two CUDA streams and overlap on the GPU; autograd replays each branch's
        # backward on the stream its forward ran on.
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.overlap_streams else main
        side.wait_stream(main)
        with torch.cuda.stream(side):
            # The reference evaluates the bottleneck twice per domain (a2gnn.py:181 & :192, :193 & :211) on
            # identical inputs with fresh dropout masks.  Both evaluations are made in one pass
            # (A2GNNBase.feat_bottleneck_pair): layer 1 once, later layers on the stacked pair -- same values
            # (SURVEY.md Appendix B.2).
            s1 = net.first_conv(source_data.x, source_data.edge_index, self.s_pnums)
            s_feat_1, source_features = net.feat_bottleneck_pair(                         # :181 (bottleneck), :192
                source_data.x, source_data.edge_index, source_batch, self.s_pnums, first_layer=s1)
            source_logits = net.feat_classifier(s_feat_1, source_data.edge_index, source_batch, prop_nums=1)
            train_loss = ops.softmax_cross_entropy(source_logits, source_data.y)          # :182
        t1 = net.first_conv(target_data.x, target_data.edge_index, self.t_pnums)
        target_features, t_feat_2 = net.feat_bottleneck_pair(                             # :193, :211 (bottleneck)
            target_data.x, target_data.edge_index, target_batch, self.t_pnums, first_layer=t1)
        if side is not main:
            main.wait_stream(side)
            for t in (source_logits, train_loss, source_features):
                t.record_stream(main)
        if self.adv:                                                                      # :196-205
            source_dlogits = net.domain_classifier(source_features, alpha)
            target_dlogits = net.domain_classifier(target_features, alpha)
            n_s, n_t = source_data.x.shape[0], target_data.x.shape[0]
            dlogits = torch.cat([source_dlogits, target_dlogits], 0)
            if dlogits.shape[0] != n_s + n_t:
                # the reference builds node-count labels against pooled features in graph
                # mode and F.cross_entropy raises; keep that behaviour
                raise ValueError("Expected input batch_size ({}) to match target batch_size ({})."
                                 .format(dlogits.shape[0], n_s + n_t))
            domain_loss = ops.domain_cross_entropy(dlogits, n_s)
            loss = ops.combine([(train_loss, 1.0), (domain_loss, float(self.weight))])
        else:                                                                             # :207-209
            mmd_loss = MMD(source_features, target_features, indices=mmd_indices)
            loss = ops.combine([(train_loss, 1.0), (mmd_loss, float(self.weight))] +
                               self._extra_loss_terms(target_features, target_data))

        target_logits = net.feat_classifier(t_feat_2, target_data.edge_index, target_batch, prop_nums=1)   # :211
        return loss, source_logits, target_logits

    def _extra_loss_terms(self, target_features, target_data):
        """[(scalar tensor, weight), ...] added to the MMD objective by subclasses (TDSS)."""
        return []

    @staticmethod
    def alpha_at(epoch, total):
        p = float(epoch) / total                                                          # :305
        return 2. / (1. + np.exp(-10. * p)) - 1                                           # :306

    def backward_and_step(self, loss, optimizer):
        """a2gnn.py:317-319."""
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()

    def train_step(self, source_data, target_data, alpha, optimizer, mmd_indices=None):
        """The loop body at a2gnn.py:309-319; returns (loss tensor, source logits, target logits)."""
        self.a2gnn.train()
        source_data = source_data.to(self.device)
        target_data = target_data.to(self.device)
        loss, source_logits, target_logits = self.forward_model(source_data, target_data, alpha,
                                                                mmd_indices=mmd_indices)
        self.backward_and_step(loss, optimizer)
        return loss, source_logits, target_logits, source_data

    def prepare_fit(self, source_data, target_data):
        """Everything ``fit`` does before its epoch loop (loaders :254-288, model :290, optimiser :292-296); returns
        ``run_epoch(epoch, last=False) -> (loss tensor, source logits, source labels)`` for the full-batch node-level
        CUDA-graph path, or None when the eager loop over the loaders applies (graph mode, ``cuda_graph = False``)."""
        graphed = bool(self.cuda_graph) and self.mode == 'node' and self.epoch > 1 and self.batch_size == 0 \
            and str(self.device).startswith('cuda')
        # the graphed step stages the next epoch's copy itself; the loaders' own prefetch is for the eager loop
        self._build_loaders(source_data, target_data, prefetch=False if graphed else None)
        self.a2gnn = self.init_model(**self.kwargs)
        optimizer = Adam(self.a2gnn.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer = optimizer
        if not graphed:
            return None
        # The loop body replayed from a CUDA graph: epoch 0 runs eagerly inside GraphedStep (it warms the caches the
        # capture must not allocate), epochs 1.. are replays; a host-resident graph is re-sent every epoch, the copy
        # of epoch e+1 overlapping the replay of epoch e (models/graphed.py).
        from .graphed import GraphedStep
        source_batch = next(iter(self.source_loader))
        target_batch = next(iter(self.target_loader))
        step = GraphedStep(self, source_batch, target_batch, optimizer, warmup=1,
                           alpha_fn=lambda i: self.alpha_at(i, self.epoch))
        self.graphed_step = step

        def run_epoch(epoch, last=False):
            if epoch == 0:
                loss, source_logits, _ = step.warmup_results[0]
            else:
                loss, source_logits, _ = step(self.alpha_at(epoch, self.epoch), last=last)
            return loss, source_logits, step.src.y
        return run_epoch

    def fit(self, source_data, target_data):
        run_epoch = self.prepare_fit(source_data, target_data)
        if run_epoch is not None:
            return self._fit_graphed(run_epoch)
        optimizer = self.optimizer

        def step(epoch, sampled_source_data, sampled_target_data):
            alpha = self.alpha_at(epoch, self.epoch)
            loss, source_logits, _, sampled_source_data = self.train_step(
                sampled_source_data, sampled_target_data, alpha, optimizer)
            return loss, source_logits, sampled_source_data

        self._fit_loop(step)

    def _fit_graphed(self, run_epoch):
        """The epoch loop of fit() (:300-336) over the graphed step.  Logging as in _fit_loop."""
        import time
        from ..metrics import micro_f1_from_logits
        from ..utils import logger
        start_time = time.time()
        for epoch in range(self.epoch):
            loss, source_logits, source_labels = run_epoch(epoch, last=epoch == self.epoch - 1)
            epoch_loss = loss.item()
            micro_f1_score = None
            if self.verbose > 1:
                micro_f1_score = micro_f1_from_logits(source_labels, source_logits)
            logger(epoch=epoch, loss=epoch_loss, source_train_acc=micro_f1_score,
                   time=time.time() - start_time, verbose=self.verbose, train=True)

    def process_graph(self, data):
        pass

    def predict(self, data, source=False):
        self.a2gnn.eval()
        pnums = self.s_pnums if source else self.t_pnums
        return self._predict_loop(lambda d: self.a2gnn(d, pnums), source)
