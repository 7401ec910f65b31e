"""GRADE over a 1-D node partition (one process per GPU) -- BASELINE.json config 4 ("GRADE (Gaussian-MMD) ...
1-D node partition, 4 x B200").  Same model, loss and gradients as ``GRADE`` on the whole graph
(``tests/dist_check.py``).

Data: ``pygda_b200.dist.partition_data(full_data, group)`` per domain.  The stock ``GCNConv`` layers
(pygda/nn/grade_base.py:58-61) take the NVLink peer-memory aggregation through the partition tag on
``edge_index``; everything else is row-local.  Exchanged per step: one scalar all-reduce for the
cross-entropy mean (grade.py:165), for ``disc='MMD'`` one small all-reduce per domain for the sampled rows
(indices over ``[:mind]`` drawn once on rank 0 exactly as pygda/utils/mmd.py:148-149, grade.py:177-182), for
``disc='JS'`` one scalar all-reduce for the domain cross-entropy mean (grade.py:169-176), and one flat
all-reduce of the weight gradients."""
import torch
import torch.distributed as dist

from .. import ops
from ..dist import AllReduceSum, GatherRows, allreduce_grads
from ..utils import draw_indices
from ..utils.mmd import to_device_async
from .grade import GRADE


class DistGRADE(GRADE):
    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        if group is None:
            raise ValueError("DistGRADE needs a pygda_b200.dist.PeerGroup")
        if self.mode != 'node' or self.disc not in ('JS', 'MMD'):
            raise NotImplementedError("the partitioned path covers node-level GRADE with disc in {'JS', 'MMD'}")
        self.group = group

    def init_model(self, **kwargs):
        net = super().init_model(**kwargs)
        for p in net.parameters():                       # replicas start from rank 0's weights
            dist.broadcast(p.data, src=0, group=self.group.pg)
        return net

    def _mmd_indices(self, mind, given):
        if given is not None:
            s_idx, t_idx = given
        elif self.group.rank == 0:
            s_idx, t_idx = draw_indices(mind, mind)
        else:
            s_idx = t_idx = torch.empty(5, 1000, dtype=torch.int64)
        dev = self.group.device
        s_idx, t_idx = to_device_async(s_idx, dev), to_device_async(t_idx, dev)
        if given is None:
            dist.broadcast(s_idx, src=0, group=self.group.pg)
            dist.broadcast(t_idx, src=0, group=self.group.pg)
        return s_idx, t_idx

    def forward_model(self, source_data, target_data, alpha, mmd_indices=None):
        net, pg = self.grade, self.group.pg
        source_logits, source_feats = net(source_data)                                    # grade.py:163
        target_logits, target_feats = net(target_data)                                    # :164
        n_s, n_t = source_data.num_nodes_global, target_data.num_nodes_global
        ce_local = ops.softmax_cross_entropy(source_logits, source_data.y)                # :165, local mean
        train_loss = AllReduceSum.apply(ops.combine([(ce_local, source_data.x.shape[0] / float(n_s))]), pg)
        if self.disc == 'JS':                                                             # :169-176
            feats = torch.cat([source_feats, target_feats], dim=0)
            domain_preds = net.discriminator(ops.GradReverse.apply(feats, alpha))
            d_local = ops.domain_cross_entropy(domain_preds, source_feats.shape[0])
            domain_loss = AllReduceSum.apply(ops.combine([(d_local, feats.shape[0] / float(n_s + n_t))]), pg)
        else:                                                                             # :177-182
            mind = min(n_s, n_t)
            s_idx, t_idx = self._mmd_indices(mind, mmd_indices)
            times, b = s_idx.shape
            s_rows = GatherRows.apply(source_feats, s_idx.reshape(-1), source_data.row_lo, pg)
            t_rows = GatherRows.apply(target_feats, t_idx.reshape(-1), target_data.row_lo, pg)
            ar = torch.arange(times * b, device=s_rows.device).view(times, b)
            domain_loss = ops.MMDFn.apply(s_rows, t_rows, ar, ar.clone(), 2.0, 5)         # replicated on every rank
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
        allreduce_grads(list(self.grade.parameters()), self.group.pg)
        optimizer.step()
        return loss, source_logits, target_logits, source_data
