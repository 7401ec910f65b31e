"""The steps either side of the hot path on the device: graph-batch collation
(gda_collate_graphs) against Batch.from_data_list, and the fused argmax + confusion count
(gda_argmax_confusion) against torch.argmax + sklearn's f1_score, which is what the reference calls
(pygda/models/a2gnn.py:328-329, pygda/metrics/metrics.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_collation_equals_the_oracle_collation():
    """Device collation (gda_collate_graphs) and the host ``Batch.from_data_list`` against the ORACLE's restatement of
    PyG's ``Batch.from_data_list`` / ``DataLoader`` (oracle/data.py: collate_graphs, GraphDataLoader) -- bit-exact."""
    from oracle import data as OD
    from pygda_b200.data import Batch, DataLoader, DeviceGraphDataset
    from pygda_b200.synthetic import graph_dataset
    ds = graph_dataset(200, 30, 2.05, 14, 2, seed=0)
    ods = [OD.Data(x=d.x, edge_index=d.edge_index, y=d.y) for d in ds]
    res = DeviceGraphDataset(ds, "cuda:0")
    g = torch.Generator().manual_seed(1)
    for ids in ([0], [199, 0, 57], torch.randperm(200, generator=g)[:64].tolist(), list(range(200))):
        ref = OD.collate_graphs([ods[i] for i in ids])
        ptr = torch.zeros(len(ids) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor([ods[i].x.size(0) for i in ids]), 0)
        for got in (res.collate(ids), Batch.from_data_list([ds[i] for i in ids])):
            assert torch.equal(got.x.cpu(), ref.x) and torch.equal(got.edge_index.cpu(), ref.edge_index)   # bit-exact
            assert torch.equal(got.y.cpu(), ref.y) and torch.equal(got.batch.cpu(), ref.batch)
            assert torch.equal(got.ptr.cpu(), ptr) and len(got) == len(ref) == len(ids)
    # the loaders: the oracle's shuffle order and batches for the same CPU seed, host and device path
    torch.manual_seed(3)
    ora = [b for b in OD.GraphDataLoader(ods, batch_size=64, shuffle=True)]
    torch.manual_seed(3)
    host = [b for b in DataLoader(ds, batch_size=64, shuffle=True)]
    torch.manual_seed(3)
    dev = [b for b in DataLoader(ds, batch_size=64, shuffle=True, device="cuda:0")]
    assert len(ora) == len(host) == len(dev) == 4
    for o, h, d in zip(ora, host, dev):
        for b in (h, d):
            assert torch.equal(b.x.cpu(), o.x) and torch.equal(b.edge_index.cpu(), o.edge_index)
            assert torch.equal(b.y.cpu(), o.y) and torch.equal(b.batch.cpu(), o.batch) and len(b) == len(o)
        assert d.x.is_cuda


def test_full_batch_loader_equals_the_oracle_loader():
    """NeighborLoader in full-batch mode (edges regrouped by destination, node order kept, extra attributes carried)
    against oracle/data.py: FullBatchNeighborLoader -- index arrays bit-exact, on host and on device-resident input."""
    from oracle import data as OD
    from pygda_b200.data import Data, NeighborLoader
    from pygda_b200.synthetic import citation_graph
    d = citation_graph(3000, 24000, 40, 4, seed=9)
    w = torch.rand(d.edge_index.size(1))
    ref = next(iter(OD.FullBatchNeighborLoader(OD.Data(x=d.x, edge_index=d.edge_index, y=d.y, edge_weight=w,
                                                       note="kept"))))
    for dev in ("cpu", "cuda:0"):
        src = Data(x=d.x, edge_index=d.edge_index, y=d.y, edge_weight=w, note="kept").to(dev)
        got = next(iter(NeighborLoader(src, [-1, -1], batch_size=3000)))
        assert torch.equal(got.edge_index.cpu(), ref.edge_index) and torch.equal(got.edge_weight.cpu(), ref.edge_weight)
        assert torch.equal(got.x.cpu(), ref.x) and torch.equal(got.y.cpu(), ref.y) and got.note == "kept"
        assert len(got) == len(ref) == 1


@pytest.mark.parametrize("rows,c", [(1, 2), (1000, 5), (100_000, 5), (5000, 64)])
def test_confusion_counts_and_f1_match_sklearn(rows, c):
    from sklearn.metrics import confusion_matrix, f1_score
    from pygda_b200.metrics import confusion_from_logits, macro_f1_from_logits, micro_f1_from_logits
    g = torch.Generator().manual_seed(rows + c)
    logits = torch.randn(rows, c, generator=g)
    logits[::7] = logits[::7].round()                       # ties: the first maximal index wins
    labels = torch.randint(max(c - 1, 1), (rows,), generator=g)      # the last class never occurs as a label
    cm, pred = confusion_from_logits(labels.cuda(), logits.cuda(), return_pred=True)
    ref_pred = logits.argmax(dim=1)
    assert torch.equal(pred.cpu(), ref_pred)
    ref_cm = torch.from_numpy(confusion_matrix(labels.numpy(), ref_pred.numpy(), labels=list(range(c))))
    assert torch.equal(cm, ref_cm)
    assert abs(micro_f1_from_logits(labels.cuda(), logits.cuda())
               - f1_score(labels.numpy(), ref_pred.numpy(), average="micro")) < 1e-12
    assert abs(macro_f1_from_logits(labels.cuda(), logits.cuda())
               - f1_score(labels.numpy(), ref_pred.numpy(), average="macro")) < 1e-12


def test_bad_label_is_reported():
    from pygda_b200.metrics import confusion_from_logits
    with pytest.raises(ValueError):
        confusion_from_logits(torch.tensor([0, 7]).cuda(), torch.randn(2, 5).cuda())


def test_graph_mode_fit_uses_resident_dataset():
    from pygda_b200.data import DeviceGraphDataset
    from pygda_b200.models import AdaGCN
    from pygda_b200.synthetic import graph_dataset
    src = graph_dataset(96, 30, 2.05, 14, 2, seed=0)
    tgt = graph_dataset(96, 39, 3.7, 14, 2, seed=1)
    torch.manual_seed(0)
    model = AdaGCN(in_dim=14, hid_dim=32, num_classes=2, mode="graph", num_layers=2, lr=0.01, epoch=2, batch_size=32,
                   device="cuda:0", verbose=2)
    model.fit(src, tgt)
    assert isinstance(model.source_loader.resident, DeviceGraphDataset)
    logits, labels = model.predict(None)
    assert logits.shape == (96, 2) and labels.shape == (96,) and torch.isfinite(logits).all()


@pytest.mark.parametrize("n,f,density", [(3000, 6775, 0.07), (500, 40000, 0.01), (300, 70000, 0.01), (64, 33, 0.2)])
def test_packed_pinned_features_rebuild_bit_exactly(n, f, density):
    """Data.pin_memory() keeps a sparse x row-compressed; .to(cuda) must give back the same bits."""
    from pygda_b200.data import Data
    g = torch.Generator().manual_seed(n + f)
    x = torch.where(torch.rand(n, f, generator=g) < density, torch.randn(n, f, generator=g), torch.zeros(()))
    x[0, f - 1] = 1.5                                    # last column (uint16 ids above 32767 when f > 32768)
    x[1].zero_()                                         # an all-zero row
    x[2, 3] = -0.0                                       # a negative zero is data too
    d = Data(x=x, edge_index=torch.randint(n, (2, 50)), y=torch.randint(5, (n,)))
    p = d.pin_memory()
    assert "_packed_x" in p.__dict__ and p.x is x
    assert p.h2d_nbytes() < 0.5 * d.pin_memory(pack=False).h2d_nbytes()
    on = p.to("cuda:0")
    assert "_packed_x" not in on.__dict__ and on.x.is_cuda
    assert torch.equal(on.x.cpu().view(torch.int32), x.view(torch.int32))
    assert torch.equal(on.edge_index.cpu(), d.edge_index) and torch.equal(on.y.cpu(), d.y)
    again = p.to("cuda:0")
    assert torch.equal(again.x, on.x)


def test_dense_features_are_not_packed():
    from pygda_b200.data import Data
    d = Data(x=torch.randn(200, 64), edge_index=torch.randint(200, (2, 50)), y=torch.randint(5, (200,)))
    p = d.pin_memory()
    assert "_packed_x" not in p.__dict__ and p.x.is_pinned()
    assert torch.equal(p.to("cuda:0").x.cpu(), d.x)
