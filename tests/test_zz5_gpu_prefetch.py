"""Opt-in prefetching loader (pygda_b200/data.py: NeighborLoader(prefetch_device=...)): the batches it hands out are
bit-identical to ``batch.to(device)``, epoch after epoch, and ``fit`` with ``prefetch = True`` trains to the same
weights as without."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


def test_prefetched_batches_equal_plain_copies():
    from pygda_b200.data import NeighborLoader
    from pygda_b200.synthetic import citation_graph
    d = citation_graph(3000, 24000, 300, 4, seed=3)                 # sparse x: row-compressed pinned staging
    plain = NeighborLoader(d, [-1, -1], batch_size=3000, pin=True)
    pre = NeighborLoader(d, [-1, -1], batch_size=3000, pin=True, prefetch_device="cuda:0")
    ref = next(iter(plain)).to("cuda:0")
    torch.cuda.synchronize()
    seen = []
    for epoch in range(4):
        (b,) = list(pre)
        assert b.x.is_cuda and b.edge_index.is_cuda and b.y.is_cuda
        # consume on the current stream right away, as a training step would
        assert torch.equal(b.x, ref.x) and torch.equal(b.edge_index, ref.edge_index) and torch.equal(b.y, ref.y)
        seen.append(b)
    assert len({id(b) for b in seen}) == 4                          # a new batch object per epoch
    keys = {b.edge_index._gda_key for b in seen}                    # ... that maps to ONE cached graph
    assert len(keys) == 1 and None not in keys
    assert b.to("cuda:0") is b                                      # already resident: the fit loops' .to() is a no-op


def _fit_arm(src, tgt, cuda_graph, prefetch, epochs=4):
    from pygda_b200.models import A2GNN
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    model = A2GNN(in_dim=64, hid_dim=32, num_classes=3, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=3,
                  weight=1, epoch=epochs, lr=0.01, device="cuda:0", verbose=0)
    if cuda_graph is not None:
        model.cuda_graph = cuda_graph
    if prefetch is not None:
        model.prefetch = prefetch
    model.fit(src, tgt)
    return model, model.predict(tgt)[0]


def test_fit_with_prefetch_matches_fit_without():
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(2000, 16000, 64, 3, seed=2)              # host-resident
    out = [_fit_arm(src, tgt, False, prefetch)[1] for prefetch in (False, True)]
    assert_close(out[1], out[0], 1e-4, "logits after 4 epochs, prefetch vs plain")


def test_default_fit_is_graphed_and_staged_and_matches_the_serial_eager_fit():
    """fit()'s DEFAULT on host-resident graphs: CUDA-graph replay with the next epoch's copy double-buffered behind it
    (models/graphed.py: StagedBatch) -- same trajectory as the reference's schedule (copy, then compute, eager)."""
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(2000, 16000, 64, 3, seed=2)
    _, serial = _fit_arm(src, tgt, False, False, epochs=5)
    model, default = _fit_arm(src, tgt, None, None, epochs=5)
    gs = model.graphed_step
    assert gs.replays == 4 and all(sb is not None for sb in gs.staged)
    assert gs.h2d_bytes_per_step == sum(sb.nbytes for sb in gs.staged) > 0
    assert_close(default, serial, 1e-4, "logits after 5 epochs, default (graphed + staged) vs eager serial")


def test_staged_step_consumes_the_bytes_copied_for_that_epoch():
    """Every replay trains on the host data as it was when that epoch's copy was issued: scaling the pinned host
    values in place changes the logits of the replay AFTER the one whose copy was already in flight."""
    from pygda_b200.data import NeighborLoader
    from pygda_b200.models import A2GNN
    from pygda_b200.models.graphed import GraphedStep
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1500, 12000, 300, 3, seed=5)             # sparse x: packed staging
    sb = next(iter(NeighborLoader(src, [-1, -1], batch_size=1500, pin=True)))
    tb = next(iter(NeighborLoader(tgt, [-1, -1], batch_size=1500, pin=True)))
    assert "_packed_x" in sb.__dict__
    est = A2GNN(in_dim=300, hid_dim=32, num_classes=3, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=2, weight=1,
                epoch=50, lr=0.0, device="cuda:0", verbose=0)       # lr = 0: the weights never move
    torch.manual_seed(0)
    est.a2gnn = est.init_model()
    opt = Adam(est.a2gnn.parameters(), lr=0.0)
    g = GraphedStep(est, sb, tb, opt, warmup=1)
    base = g()[1].clone()                                           # consumes copy #1, issues copy #2
    again = g()[1].clone()                                          # consumes copy #2 (issued before the edit below)
    assert_close(again, base, 1e-6, "unchanged inputs")
    torch.cuda.synchronize()
    sb.__dict__["_packed_x"].vals.mul_(2.0)                         # host edit; copy #3 is already in flight
    if getattr(sb.__dict__["_packed_x"], "compressed", False):      # exponent-packed staging arrays: rebuilt IN PLACE
        sb.__dict__["_packed_x"].compress()                         # (same pinned buffers the staging plan copies from)
    g()                                                             # consumes #3 (may or may not see the edit), issues #4
    edited = g()[1].clone()                                         # consumes #4: issued after the edit
    assert_close(edited, 2.0 * base, 1e-5, "source logits follow the re-sent host data (linear first layer, no bias)")
