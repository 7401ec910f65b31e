"""Host-side logic of the multi-GPU path on CPU: partition arithmetic and the autograd
collectives (world_size 2, gloo).  The GPU side is checked by tests/dist_check.py on 2 B200s."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygda_b200.dist import AllReduceSum, GatherRows, allreduce_grads, block_range, pack_column, rows_per_rank


def test_block_ranges_cover_everything_once():
    for n in (0, 1, 7, 100, 100_001):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = block_range(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= rows_per_rank(n, world)
                covered += list(range(lo, hi)) if n < 1000 else []
            if n < 1000:
                assert covered == list(range(n))
            assert sum(block_range(n, world, r)[1] - block_range(n, world, r)[0] for r in range(world)) == n


def test_column_packing_roundtrip():
    rpr = rows_per_rank(100_001, 8)
    for col in (0, 1, rpr - 1, rpr, 5 * rpr + 17, 100_000):
        p = pack_column(col, rpr)
        assert (p >> 28) * rpr + (p & 0x0FFFFFFF) == col and (p >> 28) < 8


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    n, d = 11, 3
    full = torch.randn(n, d)
    w = torch.randn(d, requires_grad=True)
    lo, hi = block_range(n, world, rank)
    idx = torch.tensor([0, 10, 3, 3, 7, 5])
    # partitioned: features = local rows * w ; loss = sum(gathered rows^2) + global mean of a local term
    feats = full[lo:hi] * w
    rows = GatherRows.apply(feats, idx, lo, None)
    local_term = feats.sum() / n
    loss = (rows ** 2).sum() + AllReduceSum.apply(local_term, None)
    loss.backward()
    allreduce_grads([w])
    # reference on the whole matrix
    w2 = w.detach().clone().requires_grad_(True)
    f2 = full * w2
    ref = (f2[idx] ** 2).sum() + f2.sum() / n
    ref.backward()
    out[rank] = (float((loss - ref).abs()), float((w.grad - w2.grad).abs().max()), torch.equal(rows, f2[idx].detach()))
    dist.destroy_process_group()


def test_autograd_collectives_world_size_2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    for r in range(2):
        dl, dg, rows_ok = out[r]
        assert dl < 1e-5 and dg < 1e-5 and rows_ok
