"""AdaGCN, graph level, data-parallel over the graphs of each mini-batch (one process per GPU) --
BASELINE.json config 5 ("AdaGCN graph-level ... 50k graphs batch=512, 8 x B200"; SURVEY.md section 8e:
"shard each mini-batch's graphs across GPUs -- pure DP").  Same loss and gradients as ``AdaGCN`` on the
whole mini-batch.

Every rank encodes its contiguous share of the batch's graphs (``pygda_b200.dist.shard_batch``).  The pooled
encodings [G, hid] are tiny, so they are all-gathered (``AllGatherRows``) and the critic, its WGAN-GP
penalty (adagcn.py:387-454) and the Wasserstein term ``|mean D(s) - mean D(t)|`` (adagcn.py:175-177,
190-192) run REPLICATED on the full batch on every rank: batch means, interpolation pairs and critic updates
are exactly those of the single-GPU step, and the critic replicas stay in step without any exchange (same
inputs, same interpolation coefficients -- broadcast from rank 0 --, same dropout stream: the CUDA generator
is seeded identically on every rank).  The classifier runs on the local graphs; its cross-entropy mean is a
scalar all-reduce.  Encoder + classifier gradients: one flat all-reduce per step (11 per step in a
per-iteration DP scheme; the critic loop here needs none)."""
import torch
import torch.distributed as dist

from .. import ops
from ..dist import AllGatherRows, AllReduceSum, allreduce_grads
from .adagcn import AdaGCN


class DistAdaGCN(AdaGCN):
    def __init__(self, *args, pg=None, **kwargs):
        super().__init__(*args, **kwargs)
        if not dist.is_initialized():
            raise RuntimeError("init torch.distributed before creating a DistAdaGCN")
        if self.mode != 'graph':
            raise NotImplementedError("the data-parallel path covers graph-level AdaGCN (node level: partition)")
        self.pg = pg
        self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)

    def init_model(self, **kwargs):
        net = super().init_model(**kwargs)
        for p in net.parameters():
            dist.broadcast(p.data, src=0, group=self.pg)
        return net

    def init_critic(self):
        super().init_critic()
        for p in self.discriminator.parameters():
            dist.broadcast(p.data, src=0, group=self.pg)
        seed = torch.tensor([torch.cuda.initial_seed() & 0x7FFFFFFF], dtype=torch.int64, device=self.device)
        dist.broadcast(seed, src=0, group=self.pg)
        torch.cuda.manual_seed(int(seed.item()))          # one dropout stream for all critic replicas

    def _rand(self, n):
        a = torch.rand((n, 1)).to(self.device)            # adagcn.py:420,429,434 -- rank 0's draw, for everyone
        dist.broadcast(a, src=0, group=self.pg)
        return a

    def _encode_all(self, data):
        """Local pooled encodings -> the full batch's, in graph order, on every rank."""
        return AllGatherRows.apply(self.adagcn(data), self.pg)

    def forward_model(self, source_data, target_data):
        for _ in range(10):                                                               # adagcn.py:169-183
            with torch.no_grad():
                encoded_source = self._encode_all(source_data)
                encoded_target = self._encode_all(target_data)
            gp_loss = self.gradient_penalty(encoded_source, encoded_target)
            dis_s = torch.mean(self._critic(encoded_source).reshape(-1))
            dis_t = torch.mean(self._critic(encoded_target).reshape(-1))
            loss = - torch.abs(dis_s - dis_t) + self.gp_weight * gp_loss
            self.c_optimizer.zero_grad()
            loss.backward()
            self.c_optimizer.step()
        local_source = self.adagcn(source_data)                                           # :185-186
        local_target = self.adagcn(target_data)
        encoded_source = AllGatherRows.apply(local_source, self.pg)
        encoded_target = AllGatherRows.apply(local_target, self.pg)
        source_logits = self.adagcn.cls_model(local_source)
        g_s = encoded_source.shape[0]
        ce_local = self.adagcn.loss_func(source_logits, source_data.y)                    # local mean
        cls_loss = AllReduceSum.apply(ops.combine([(ce_local, local_source.shape[0] / float(g_s))]), self.pg)
        dis_s = torch.mean(self._critic(encoded_source).reshape(-1))
        dis_t = torch.mean(self._critic(encoded_target).reshape(-1))
        dis_loss = torch.abs(dis_s - dis_t)
        target_logits = self.adagcn.cls_model(local_target)
        loss = cls_loss + dis_loss * self.domain_weight                                   # :196
        return loss, source_logits, target_logits

    def fit(self, source_data, target_data):
        """``AdaGCN.fit`` with every mini-batch split over the ranks.  All ranks must iterate the same batches:
        seed the CPU generator identically before calling (the loaders shuffle with ``torch.randperm``)."""
        from ..dist import shard_batch
        from ..optim import Adam
        self._build_loaders(source_data, target_data)
        self.adagcn = self.init_model(**self.kwargs)
        optimizer = Adam(self.adagcn.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer = optimizer
        self.init_critic()

        def step(epoch, s, t):
            s, t = shard_batch(s, self.rank, self.world), shard_batch(t, self.rank, self.world)
            loss, source_logits, _, s = self.train_step(s, t, optimizer)
            return loss, source_logits, s

        self._fit_loop(step)

    def train_step(self, source_data, target_data, optimizer):
        self.adagcn.train()
        source_data = source_data.to(self.device)
        target_data = target_data.to(self.device)
        loss, source_logits, target_logits = self.forward_model(source_data, target_data)
        optimizer.zero_grad()
        loss.backward()
        allreduce_grads(list(self.adagcn.parameters()), self.pg)
        optimizer.step()
        return loss, source_logits, target_logits, source_data
