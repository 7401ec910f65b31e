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


def test_fit_with_prefetch_matches_fit_without():
    from pygda_b200.models import A2GNN
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(2000, 16000, 64, 3, seed=2)              # host-resident
    out = []
    for prefetch in (False, True):
        torch.manual_seed(0)
        torch.cuda.manual_seed(0)
        model = A2GNN(in_dim=64, hid_dim=32, num_classes=3, num_layers=2, dropout=0.0, s_pnums=0, t_pnums=3,
                      weight=1, epoch=4, lr=0.01, device="cuda:0", verbose=0)
        model.prefetch = prefetch
        model.fit(src, tgt)
        logits, _ = model.predict(tgt)
        out.append(logits)
    assert_close(out[1], out[0], 1e-4, "logits after 4 epochs, prefetch vs plain")
