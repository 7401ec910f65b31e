#!/usr/bin/env python
"""Multi-GPU parity: DistA2GNN over P ranks == A2GNN on the whole graph (same weights, dropout 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pygda_b200 import ops
    from pygda_b200.dist import PartitionedGraph, PeerGroup, partition_data
    from pygda_b200.graph import Graph
    from pygda_b200.models import A2GNN
    from pygda_b200.models.dist_a2gnn import DistA2GNN
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import domain_pair
    from pygda_b200.utils import draw_indices
    group = PeerGroup()
    dev = group.device
    ok = True

    # ---- 1. peer aggregation == single-GPU aggregation on the owned rows -----------------------
    n, h = 30011, 128                                       # not divisible by the world size
    src, tgt = domain_pair(n, 300000, 64, 5, seed=3, target_nodes=n - 1000, target_edges=280000)
    full = Graph(src.edge_index.to(dev), n)
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(1)).to(dev)
    lo, hi = group.block(n)
    bias = torch.randn(h, generator=torch.Generator().manual_seed(2)).to(dev)
    for mode in ("peer", "push", "halo"):    # in-kernel NVLink gathers / producer-side full push / halo exchange (dist.py)
        os.environ["GDA_DIST_MODE"] = mode
        part = PartitionedGraph(group, src.edge_index.to(dev), n)
        assert part.mode == mode
        for k, transpose in ((1, False), (3, False), (2, True), (4, False)):
            ref = ops.spmm_k(full, x, k, transpose=transpose, bias=bias, relu=True)[lo:hi]
            out = part.spmm_k(x[lo:hi].contiguous(), k, transpose=transpose, bias=bias, relu=True)
            e = rel(out, ref)
            ok &= e < 1e-6
            again = part.spmm_k(x[lo:hi].contiguous(), k, transpose=transpose, bias=bias, relu=True)
            ok &= bool(torch.equal(out, again))
            if rank == 0:
                print(f"{mode} spmm k={k} T={transpose}: rel err {e:.2e}, remote fraction {part.remote_fraction:.2f}")
        if mode == "halo":
            # two stacked matrices per exchange (the paired bottleneck evaluations): each equals its own single pass
            x2 = torch.cat([x[lo:hi], x[lo:hi].flip(1)]).contiguous()
            for k, transpose in ((1, False), (3, False), (4, True)):
                both = part.spmm_k(x2, k, transpose=transpose, bias=bias, relu=True, nb=2)
                one_a = part.spmm_k(x2[:hi - lo].contiguous(), k, transpose=transpose, bias=bias, relu=True)
                one_b = part.spmm_k(x2[hi - lo:].contiguous(), k, transpose=transpose, bias=bias, relu=True)
                same = bool(torch.equal(both[:hi - lo], one_a)) and bool(torch.equal(both[hi - lo:], one_b))
                ok &= same
                if rank == 0:
                    print(f"halo spmm nb=2 k={k} T={transpose}: identical to two single passes: {same}, "
                          f"factored (unit-weight) chain: {part._halo.unit}")
        group.check()
        del part
    os.environ.pop("GDA_DIST_MODE", None)

    # ---- 2. one training step: loss, logits, updated weights ------------------------------------
    hp = dict(in_dim=64, hid_dim=32, num_classes=5, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=4, weight=10,
              weight_decay=0.005, lr=0.01, epoch=200, verbose=0)
    torch.manual_seed(0)
    single = A2GNN(device=str(dev), **hp)
    single.a2gnn = single.init_model()
    single.overlap_streams = False
    multi = DistA2GNN(device=str(dev), group=group, **hp)
    multi.a2gnn = multi.init_model()
    multi.a2gnn.load_state_dict(single.a2gnn.state_dict())
    for p in multi.a2gnn.parameters():
        dist.broadcast(p.data, src=0)
    single.a2gnn.load_state_dict(multi.a2gnn.state_dict())
    torch.manual_seed(5)
    idx = draw_indices(n, n - 1000)
    idx = tuple(t.to(dev) for t in idx)
    for t in idx:
        dist.broadcast(t, src=0)
    idx = tuple(t.cpu() for t in idx)
    s_full, t_full = src.to(dev), tgt.to(dev)
    s_part, t_part = partition_data(src, group), partition_data(tgt, group)
    o1 = Adam(single.a2gnn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    o2 = Adam(multi.a2gnn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    for step in range(2):
        l1, sl1, tl1, _ = single.train_step(s_full, t_full, 0.3, o1, mmd_indices=idx)
        l2, sl2, tl2, _ = multi.train_step(s_part, t_part, 0.3, o2, mmd_indices=idx)
        slo, shi = group.block(n)
        tlo, thi = group.block(n - 1000)
        errs = {"loss": rel(l2, l1), "source_logits": rel(sl2, sl1[slo:shi]), "target_logits": rel(tl2, tl1[tlo:thi])}
        for (k, p), (_, q) in zip(multi.a2gnn.named_parameters(), single.a2gnn.named_parameters()):
            errs["param " + k] = rel(p, q)
        worst = max(errs.values())
        ok &= worst < 2e-4
        if rank == 0:
            print(f"step {step}: " + ", ".join(f"{k}={v:.1e}" for k, v in errs.items() if not k.startswith("param")),
                  f"max param err {max(v for k, v in errs.items() if k.startswith('param')):.1e}")
    group.check()

    # ---- 2b. the partitioned step captured as a CUDA graph (fit()'s default) == the single-GPU eager steps -----
    from pygda_b200.models.graphed import GraphedStep
    torch.manual_seed(0)
    single = A2GNN(device=str(dev), **hp)
    single.a2gnn = single.init_model()
    single.overlap_streams = False
    multi = DistA2GNN(device=str(dev), group=group, **hp)
    multi.a2gnn = multi.init_model()
    single.a2gnn.load_state_dict(multi.a2gnn.state_dict())
    o1 = Adam(single.a2gnn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    o2 = Adam(multi.a2gnn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    torch.manual_seed(77)
    d0 = draw_indices(n, n - 1000)                            # what rank 0 draws inside the warm-up step below
    torch.manual_seed(77)
    gs = GraphedStep(multi, s_part, t_part, o2, warmup=1,     # one eager step inside, on rank 0's draw (broadcast)
                     alpha_fn=lambda i: 0.3)
    l1, sl1, tl1, _ = single.train_step(s_full, t_full, 0.3, o1, mmd_indices=d0)
    l2, sl2, tl2 = gs.warmup_results[0]
    torch.manual_seed(21)
    for step in range(4):
        if step > 0:
            d = tuple(t.to(dev) for t in draw_indices(n, n - 1000))
            for t in d:
                dist.broadcast(t, src=0)
            d = tuple(t.cpu() for t in d)
            l1, sl1, tl1, _ = single.train_step(s_full, t_full, 0.3, o1, mmd_indices=d)
            l2, sl2, tl2 = gs(mmd_indices=d)
        errs = {"loss": rel(l2, l1), "source_logits": rel(sl2, sl1[slo:shi]), "target_logits": rel(tl2, tl1[tlo:thi])}
        perr = max(rel(p, q) for p, q in zip(multi.a2gnn.parameters(), single.a2gnn.parameters()))
        # forward quantities tight; weights at the bar of tests/test_gpu_graphed.py (Adam divides by sqrt(v): last-bit
        # differences of atomically accumulated reductions are amplified where a gradient element is ~0)
        ok &= max(errs.values()) < 1e-4 and perr < 1e-3
        if rank == 0:
            print(f"graphed dist step {step}{' (eager warm-up)' if step == 0 else ''}: " +
                  ", ".join(f"{k}={v:.1e}" for k, v in errs.items()),
                  f"max param err {perr:.1e}, {gs.launches_per_replay} kernels per replay")
    group.check()
    del gs

    # ---- 3. GRADE over the partition (config 4): MMD and JS discrepancies ------------------------
    from pygda_b200.models import GRADE
    from pygda_b200.models.dist_grade import DistGRADE
    for disc in ("MMD", "JS"):
        hp = dict(in_dim=64, hid_dim=32, num_classes=5, num_layers=2, dropout=0.0, disc=disc, weight=0.5,
                  weight_decay=0.01, lr=0.01, epoch=200, verbose=0)
        torch.manual_seed(1)
        single = GRADE(device=str(dev), **hp)
        single.grade = single.init_model()
        multi = DistGRADE(device=str(dev), group=group, **hp)
        multi.grade = multi.init_model()
        single.grade.load_state_dict(multi.grade.state_dict())
        mind = n - 1000
        torch.manual_seed(6)
        idx = tuple(t.to(dev) for t in draw_indices(mind, mind))
        for t in idx:
            dist.broadcast(t, src=0)
        idx = tuple(t.cpu() for t in idx)
        o1 = Adam(single.grade.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
        o2 = Adam(multi.grade.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
        for step in range(2):
            l1, sl1, tl1, _ = single.train_step(s_full, t_full, 0.3, o1, mmd_indices=idx)
            l2, sl2, tl2, _ = multi.train_step(s_part, t_part, 0.3, o2, mmd_indices=idx)
            slo, shi = group.block(n)
            tlo, thi = group.block(n - 1000)
            errs = {"loss": rel(l2, l1), "source_logits": rel(sl2, sl1[slo:shi]), "target_logits": rel(tl2, tl1[tlo:thi])}
            perr = max(rel(p, q) for p, q in zip(multi.grade.parameters(), single.grade.parameters()))
            ok &= max(max(errs.values()), perr) < 2e-4
            if rank == 0:
                print(f"GRADE[{disc}] step {step}: " + ", ".join(f"{k}={v:.1e}" for k, v in errs.items()),
                      f"max param err {perr:.1e}")
    group.check()

    # ---- 4. graph-level AdaGCN, data-parallel over the graphs of a mini-batch (config 5) ---------
    from pygda_b200.data import Batch
    from pygda_b200.dist import shard_batch
    from pygda_b200.models import AdaGCN
    from pygda_b200.models.dist_adagcn import DistAdaGCN
    from pygda_b200.synthetic import graph_dataset
    hp = dict(in_dim=14, hid_dim=32, num_classes=2, mode="graph", num_layers=2, gp_weight=5, domain_weight=0.1,
              weight_decay=0.01, lr=0.01, epoch=2, batch_size=64, verbose=0)
    ds_s = graph_dataset(61, 30, 2.05, 14, 2, seed=0)
    ds_t = graph_dataset(64, 39, 3.7, 14, 2, seed=1)
    bs, bt = Batch.from_data_list(list(ds_s)).to(dev), Batch.from_data_list(list(ds_t)).to(dev)
    torch.manual_seed(2)
    single = AdaGCN(device=str(dev), **hp)
    single.adagcn = single.init_model()
    multi = DistAdaGCN(device=str(dev), **hp)
    multi.adagcn = multi.init_model()
    single.adagcn.load_state_dict(multi.adagcn.state_dict())
    for est in (single, multi):
        est.adagcn.encoder.dropout.p = 0.0                  # element-indexed masks differ between shards
    multi.init_critic()
    single.init_critic()
    single.discriminator.load_state_dict(multi.discriminator.state_dict())
    o1 = Adam(single.adagcn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    o2 = Adam(multi.adagcn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    ls, lt = shard_batch(bs, rank, world), shard_batch(bt, rank, world)
    glo, ghi = group.block(61)
    for step in range(2):
        torch.manual_seed(100 + step); torch.cuda.manual_seed(100 + step)
        l1, sl1, _, _ = single.train_step(bs, bt, o1)
        torch.manual_seed(100 + step); torch.cuda.manual_seed(100 + step)
        l2, sl2, _, _ = multi.train_step(ls, lt, o2)
        errs = {"loss": rel(l2, l1), "source_logits": rel(sl2, sl1[glo:ghi]),
                "encoder": max(rel(p, q) for p, q in zip(multi.adagcn.parameters(), single.adagcn.parameters())),
                "critic": max(rel(p, q) for p, q in zip(multi.discriminator.parameters(),
                                                         single.discriminator.parameters()))}
        ok &= max(errs.values()) < 5e-4
        if rank == 0:
            print(f"AdaGCN graph DP step {step}: " + ", ".join(f"{k}={v:.1e}" for k, v in errs.items()))

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(flag.item()) else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
