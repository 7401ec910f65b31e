"""A2GNN over a 1-D node partition (one process per GPU): same model, same loss, same
gradients as ``A2GNN`` on the whole graph -- ``tests/dist_check.py`` checks that on 2 GPUs.

Data: ``pygda_b200.dist.partition_data(full_data, group)`` per domain.  What is exchanged:
neighbour rows inside the aggregation kernel (NVLink peer loads), one scalar all-reduce for the
cross-entropy mean, one small all-reduce per domain for the MMD sample rows (indices drawn once
on rank 0 exactly as pygda/utils/mmd.py:148-149 and broadcast), one flat all-reduce of the
weight gradients."""
import torch
import torch.distributed as dist

from .. import ops
from ..dist import AllReduceSum, GatherRows, allreduce_grads
from ..utils import draw_indices
from ..utils.mmd import to_device_async
from .a2gnn import A2GNN


class DistA2GNN(A2GNN):
    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        if group is None:
            raise ValueError("DistA2GNN needs a pygda_b200.dist.PeerGroup")
        if self.mode != 'node' or self.adv:
            raise NotImplementedError("the partitioned path covers node-level A2GNN with the MMD loss")
        self.group = group
        self.overlap_streams = True
        self.pair_evaluations = True

    def init_model(self, **kwargs):
        net = super().init_model(**kwargs)
        for p in net.parameters():                       # replicas start from rank 0's weights
            dist.broadcast(p.data, src=0, group=self.group.pg)
        return net

    def _mmd_indices(self, n_s, n_t, given):
        if given is not None:
            s_idx, t_idx = given
        elif self.group.rank == 0:
            s_idx, t_idx = draw_indices(n_s, n_t)
        else:
            s_idx = t_idx = torch.empty(5, 1000, dtype=torch.int64)
        dev = self.group.device
        s_idx, t_idx = to_device_async(s_idx, dev), to_device_async(t_idx, dev)
        if given is None:
            dist.broadcast(s_idx, src=0, group=self.group.pg)
            dist.broadcast(t_idx, src=0, group=self.group.pg)
        return s_idx, t_idx

    def forward_model(self, source_data, target_data, alpha, mmd_indices=None):
        net, pg = self.a2gnn, self.group.pg
        # source branch (dense, no propagation steps but the classifier's) on a side stream, target
        # branch (k peer-memory propagation steps per conv) on the current one -- as in A2GNN.forward_model;
        # each partitioned graph has its own barrier channel, so the two streams never interlock
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.overlap_streams else main
        side.wait_stream(main)
        # the two bottleneck evaluations per domain (a2gnn.py:181/192, 193/211) in one stacked pass, as on one GPU
        # (A2GNNBase.feat_bottleneck_pair), when the partition can aggregate two matrices per exchange (halo mode)
        def pair_ok(data, k):
            part = getattr(data.edge_index, "_gda_partition", None)
            if part is None or not self.pair_evaluations:
                return False
            gr = part.graph(data.edge_index, 1 | 4)
            return k == 0 or gr.supports_nb(self.hid_dim, 2)
        with torch.cuda.stream(side):
            s1 = net.first_conv(source_data.x, source_data.edge_index, self.s_pnums)
            if pair_ok(source_data, self.s_pnums):
                s_feat_1, source_features = net.feat_bottleneck_pair(
                    source_data.x, source_data.edge_index, None, self.s_pnums, first_layer=s1)
                source_logits = net.feat_classifier(s_feat_1, source_data.edge_index, None, prop_nums=1)
            else:
                source_logits = net(source_data, self.s_pnums, first_layer=s1)
                source_features = net.feat_bottleneck(source_data.x, source_data.edge_index, None, self.s_pnums,
                                                      first_layer=s1)
            ce_local = ops.softmax_cross_entropy(source_logits, source_data.y)
            frac = source_data.x.shape[0] / float(source_data.num_nodes_global)
            train_loss = AllReduceSum.apply(ops.combine([(ce_local, frac)]), pg)   # global mean CE
        t1 = net.first_conv(target_data.x, target_data.edge_index, self.t_pnums)
        t_pair = pair_ok(target_data, self.t_pnums)
        if t_pair:
            target_features, t_feat_2 = net.feat_bottleneck_pair(
                target_data.x, target_data.edge_index, None, self.t_pnums, first_layer=t1)
        else:
            target_features = net.feat_bottleneck(target_data.x, target_data.edge_index, None, self.t_pnums,
                                                  first_layer=t1)
        if side is not main:
            main.wait_stream(side)
            for t in (source_logits, train_loss, source_features):
                t.record_stream(main)
        s_idx, t_idx = self._mmd_indices(source_data.num_nodes_global, target_data.num_nodes_global, mmd_indices)
        times, b = s_idx.shape
        s_rows = GatherRows.apply(source_features, s_idx.reshape(-1), source_data.row_lo, pg)
        t_rows = GatherRows.apply(target_features, t_idx.reshape(-1), target_data.row_lo, pg)
        ar = torch.arange(times * b, device=s_rows.device).view(times, b)
        mmd_loss = ops.MMDFn.apply(s_rows, t_rows, ar, ar.clone(), 2.0, 5)          # replicated on every rank
        loss = ops.combine([(train_loss, 1.0), (mmd_loss, float(self.weight))])
        if t_pair:
            target_logits = net.feat_classifier(t_feat_2, target_data.edge_index, None, prop_nums=1)
        else:
            target_logits = net(target_data, self.t_pnums, first_layer=t1)
        return loss, source_logits, target_logits

    def backward_and_step(self, loss, optimizer):
        optimizer.zero_grad()
        loss.backward()
        allreduce_grads(list(self.a2gnn.parameters()), self.group.pg)
        optimizer.step()
