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


def _gather_worker(rank, world, port, out):
    from pygda_b200.dist import AllGatherRows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    full = torch.randn(7, 3)
    w = torch.randn(3, requires_grad=True)
    lo, hi = block_range(7, world, rank)                       # blocks of 4 and 3 rows
    enc = AllGatherRows.apply(full[lo:hi] * w, None)
    loss = (enc.mean(0) ** 2).sum() + enc[2].sum() * enc[5].sum()     # replicated on every rank, couples the blocks
    loss.backward()
    allreduce_grads([w])
    w2 = w.detach().clone().requires_grad_(True)
    e2 = full * w2
    ref = (e2.mean(0) ** 2).sum() + e2[2].sum() * e2[5].sum()
    ref.backward()
    out[rank] = (torch.equal(enc.detach(), e2.detach()), float((loss - ref).abs()), float((w.grad - w2.grad).abs().max()))
    dist.destroy_process_group()


def test_all_gather_rows_world_size_2():
    """The replicated-loss contract of the data-parallel graph-level path (models/dist_adagcn.py)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(2, 29537, out), nprocs=2, join=True)
    for r in range(2):
        same, dl, dg = out[r]
        assert same and dl < 1e-5 and dg < 1e-5


def test_shard_batch_splits_graphs_contiguously():
    from pygda_b200.data import Batch
    from pygda_b200.dist import shard_batch
    from pygda_b200.synthetic import graph_dataset
    ds = graph_dataset(11, 6, 2.0, 4, 2, seed=3)
    full = Batch.from_data_list(ds)
    for world in (1, 2, 3, 4):
        got_x, got_y, edges = [], [], 0
        for rank in range(world):
            sh = shard_batch(full, rank, world)
            lo, hi = block_range(11, world, rank)
            assert len(sh) == hi - lo
            ref = Batch.from_data_list(ds[lo:hi]) if hi > lo else None
            if ref is not None:
                assert torch.equal(sh.x, ref.x) and torch.equal(sh.y, ref.y) and torch.equal(sh.batch, ref.batch)
                assert torch.equal(sh.edge_index, ref.edge_index) and torch.equal(sh.ptr, ref.ptr)
            got_x.append(sh.x); got_y.append(sh.y); edges += sh.edge_index.size(1)
        assert torch.equal(torch.cat(got_x), full.x) and torch.equal(torch.cat(got_y), full.y)
        assert edges == full.edge_index.size(1)


def test_packed_rows_format_roundtrip_on_the_host():
    """Host half of the row-compressed pinned staging (pygda_b200/data.py: PackedRows): the values, the delta-coded
    column ids (one byte per entry, 255 = escape that only advances) and the two row-pointer arrays reproduce the
    matrix bit for bit under the decoding rule gda_unpack_rows_delta_f32 implements on the device."""
    import numpy as np
    from pygda_b200.data import Data, PackedRows
    g = torch.Generator().manual_seed(0)
    for n, f in ((300, 6775), (40, 70000), (5, 1)):
        x = torch.where(torch.rand(n, f, generator=g) < 0.05, torch.randn(n, f, generator=g), torch.zeros(()))
        x[0, f - 1] = 2.0
        x[1].zero_()                                          # an empty row
        if f > 3:
            x[2, 3] = -0.0                                    # kept: the BIT PATTERN is non-zero
        if f > 2000:
            x[3].zero_()
            x[3, 700], x[3, 1465], x[3, f - 1] = 3.0, 4.0, 5.0      # a gap of exactly 3 * 255 and a long run of escapes
        p = PackedRows(x, chunk=64)
        assert p.shape == (n, f)
        vals, dl = p.vals.numpy(), p.deltas.numpy()
        vp, bp = p.val_ptr.numpy().astype(np.int64), p.byte_ptr.numpy().astype(np.int64)
        dense = np.zeros((n, f), dtype=np.float32)
        for r in range(n):
            col, k = 0, 0
            for byte in dl[bp[r]:bp[r + 1]]:
                col += int(byte)
                if byte != 255:
                    dense[r, col] = vals[vp[r] + k]
                    k += 1
            assert k == vp[r + 1] - vp[r]
        assert np.array_equal(dense.view(np.int32), x.numpy().view(np.int32))
        assert int(vp[-1]) == p.vals.numel() and int(bp[-1]) == p.deltas.numel()
        assert p.nbytes == 4 * p.vals.numel() + p.deltas.numel() + 2 * 4 * (n + 1)
    d = Data(x=torch.zeros(10, 8), edge_index=torch.zeros(2, 0, dtype=torch.long), y=torch.zeros(10, dtype=torch.long))
    assert d.h2d_nbytes() == 10 * 8 * 4 + 10 * 8


def test_packed_tiles_format_roundtrip_on_the_host():
    """Host half of the TILE-PACKED form (pygda_b200/data.py: PackedTiles) -- the operand form of the first layer's
    tensor-core GEMMs (csrc/gemm_xt.cu) and the pinned staging form of a sparse x: sub-tiles of 32 rows x 64 columns in
    strip-major order, entries sorted by position, one low-position byte per entry, eight cumulative segment counts per
    sub-tile.  Decoded with the rule the kernels implement (entry i sits at ((# segment boundaries <= i) << 8) | code),
    the arrays reproduce the matrix bit for bit; the strip count is padded to a multiple of 4."""
    import numpy as np
    from pygda_b200.data import Data, PackedTiles
    g = torch.Generator().manual_seed(0)
    for n, f, dens in ((300, 6775, 0.05), (33, 70, 0.5), (5, 1, 1.0), (129, 4100, 0.0005), (64, 64, 1.0)):
        x = torch.where(torch.rand(n, f, generator=g) < dens, torch.randn(n, f, generator=g), torch.zeros(()))
        x[0, f - 1] = 2.0
        if n > 1:
            x[1].zero_()                                      # an empty row
        if f > 3:
            x[2 % n, 3] = -0.0                                # kept: the BIT PATTERN is non-zero
        p = PackedTiles(x, chunk=64, pin=False)
        nkb, nstrips = -(-f // 64), -(-(-(-n // 32)) // 4) * 4
        assert p.shape == (n, f) and p.ptr.numel() == nstrips * nkb + 1 and tuple(p.seg.shape) == (nstrips * nkb, 8)
        vals, codes = p.vals.numpy(), p.codes.numpy()
        ptr, seg = p.ptr.numpy().astype(np.int64), p.seg.numpy().astype(np.int64)
        assert int(ptr[-1]) == p.vals.numel() == p.codes.numel() == int((x.view(torch.int32) != 0).sum())
        assert np.all(np.diff(seg, axis=1) >= 0) and np.array_equal(seg[:, 7], np.diff(ptr))
        dense = np.zeros((n, f), dtype=np.float32)
        for t in range(nstrips * nkb):
            last = -1
            for i in range(ptr[t + 1] - ptr[t]):
                pos = int((seg[t, :7] <= i).sum()) * 256 + int(codes[ptr[t] + i])
                assert pos > last and pos < 2048                  # strictly ascending inside a sub-tile
                last = pos
                dense[(t // nkb) * 32 + pos // 64, (t % nkb) * 64 + pos % 64] = vals[ptr[t] + i]
        assert np.array_equal(dense.view(np.int32), x.numpy().view(np.int32))
        assert torch.equal(p.decode().view(torch.int32), x.view(torch.int32))
        assert p.nbytes == 5 * p.vals.numel() + 4 * (nstrips * nkb + 1) + 16 * nstrips * nkb
        # the exponent-packed form of the values (what the pinned staging copy carries): lossless for any bit pattern
        x[0, 0] = float("inf")
        if f > 5:
            x[0, 1], x[0, 2], x[0, 4], x[0, 5] = float("nan"), 1e-42, -3.5e20, -1e-4
        q = PackedTiles(x, chunk=64, pin=False, compress_values=True)
        assert "_vals" not in q.tensors() and q.m24.numel() == 12 * -(-q.vals.numel() // 4)
        assert q.ecode.numel() == 2 * -(-q.vals.numel() // 4) and int(q.vmeta[1]) >= 1
        assert torch.equal(q.decompress_values().view(torch.int32), q.vals.view(torch.int32))
    d = Data(x=torch.zeros(10, 8), edge_index=torch.zeros(2, 0, dtype=torch.long), y=torch.zeros(10, dtype=torch.long))
    assert d.h2d_nbytes() == 10 * 8 * 4 + 10 * 8


def test_pin_memory_stages_a_sparse_x_tile_packed_and_counts_its_bytes():
    """Data.pin_memory(): a sparse fp32 x is kept tile-packed with exponent-packed values in the staging area (the
    caller's tensor stays ``.x``), a dense one is pinned as it is; h2d_nbytes() counts what ``.to(cuda)`` would copy."""
    from pygda_b200.data import Data, PackedRows, PackedTiles
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn(200, 900, generator=g) - 1.5)
    x = x / x.sum(1, keepdim=True).clamp(min=1e-9)
    ei = torch.randint(0, 200, (2, 700), generator=g)
    d = Data(x=x, edge_index=ei, y=torch.zeros(200, dtype=torch.long))
    p = d.pin_memory()
    packed = p.__dict__["_packed_x"]
    assert isinstance(packed, PackedTiles) and packed.compressed and p.x is x
    assert torch.equal(packed.decode().view(torch.int32), x.view(torch.int32))
    assert torch.equal(packed.decompress_values().view(torch.int32), packed.vals.view(torch.int32))
    other = ei.numel() * 8 + 200 * 8
    assert p.h2d_nbytes() == packed.nbytes + other
    nnz = int((x != 0).sum())
    assert packed.nbytes < 4.8 * nnz + 20 * packed.seg.shape[0] + 64          # 3.5 B value + 1 B position (+ escapes)
    assert isinstance(d.pin_memory(pack="rows").__dict__["_packed_x"], PackedRows)
    dense = Data(x=torch.randn(50, 40, generator=g), edge_index=ei[:, :10] % 50, y=torch.zeros(50, dtype=torch.long))
    assert "_packed_x" not in dense.pin_memory().__dict__
    assert "_packed_x" not in d.pin_memory(pack=False).__dict__


def test_graph_loader_batches_follow_torchs_own_sampler():
    """PyG's DataLoader is torch's DataLoader with a graph collate: for the same torch seed our loader must yield the
    same graphs per batch and consume the CPU generator identically (pygda/models/a2gnn.py:276-286)."""
    import torch.utils.data as tud
    from pygda_b200.data import Data, DataLoader
    ds = [Data(x=torch.full((2, 3), float(i)), edge_index=torch.tensor([[0], [1]]), y=torch.tensor([i])) for i in range(23)]
    for bs in (1, 5, 23, 40):
        torch.manual_seed(5)
        ours = [b.y.tolist() for b in DataLoader(ds, batch_size=bs, shuffle=True)]
        after_ours = torch.rand(1)
        torch.manual_seed(5)
        ref = [list(map(int, b)) for b in tud.DataLoader(list(range(23)), batch_size=bs, shuffle=True, collate_fn=list)]
        assert ours == ref and torch.equal(after_ours, torch.rand(1))
        assert len(DataLoader(ds, batch_size=bs)) == len(ref)
    plain = [b.y.tolist() for b in DataLoader(ds, batch_size=10, shuffle=False)]
    assert plain == [list(range(10)), list(range(10, 20)), [20, 21, 22]]
    b = next(iter(DataLoader(ds, batch_size=4, shuffle=False)))
    assert len(b) == 4 and b.batch.tolist() == [0, 0, 1, 1, 2, 2, 3, 3] and b.edge_index.tolist() == [[0, 2, 4, 6], [1, 3, 5, 7]]
