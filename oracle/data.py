"""Oracle restatement of the PyG containers/loaders the reference's ``fit``
uses (pygda/models/a2gnn.py:254-288).

TEST INFRASTRUCTURE (see oracle/__init__.py).  "parity unpinned": PyG is not
installable here, behaviour restated from torch_geometric 2.4.x.
"""
import torch


class Data:
    """Duck-typed stand-in for torch_geometric.data.Data."""

    def __init__(self, x=None, edge_index=None, y=None, batch=None, num_graphs=None, **kw):
        self.x, self.edge_index, self.y, self.batch = x, edge_index, y, batch
        self.num_graphs = num_graphs
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        out = Data.__new__(Data)
        for k, v in self.__dict__.items():
            setattr(out, k, v.to(device) if torch.is_tensor(v) else v)
        return out

    @property
    def num_nodes(self):
        return self.x.size(0)

    def __len__(self):
        # PyG ``Batch.__len__`` == number of graphs (pygda/models/a2gnn.py:270-271)
        return self.num_graphs if self.num_graphs is not None else 1


class FullBatchNeighborLoader:
    """``NeighborLoader(data, [-1]*L, batch_size=N)`` in full-batch mode: one
    batch per epoch holding the whole graph, node order preserved, every in-edge
    once, edges regrouped by destination (stable).  Appendix A.6."""

    def __init__(self, data, num_neighbors=None, batch_size=None):
        self.data = data
        ei = data.edge_index
        order = torch.argsort(ei[1], stable=True)
        # attributes PyG classifies as neither node- nor edge-level (e.g. TDSS's edge_index_smooth,
        # pygda/models/tdss.py:503) are copied through unchanged by NeighborLoader's filter_data
        extra = {k: v for k, v in data.__dict__.items() if k not in ("x", "edge_index", "y", "batch", "num_graphs")}
        for k in ("edge_weight", "edge_attr"):      # edge-level attributes are filtered with their edges
            v = extra.get(k)
            if torch.is_tensor(v) and v.dim() >= 1 and v.size(0) == ei.size(1):
                extra[k] = v[order].contiguous()
        self._batch = Data(x=data.x, edge_index=ei[:, order].contiguous(), y=data.y,
                           batch=getattr(data, "batch", None), **extra)

    def __iter__(self):
        # a NEW Data per batch, as PyG's filter_data builds one: attributes set on a yielded batch
        # (pygda/models/strurw.py:487) are gone the next epoch
        out = Data.__new__(Data)
        out.__dict__.update(self._batch.__dict__)
        yield out

    def __len__(self):
        return 1


def collate_graphs(graphs):
    """PyG ``Batch.from_data_list``: node-wise cat of x, edge_index offset by the
    running node count, y cat, batch vector = graph id per node."""
    xs, eis, ys, bs = [], [], [], []
    off = 0
    for g, d in enumerate(graphs):
        xs.append(d.x)
        eis.append(d.edge_index + off)
        ys.append(d.y.view(-1))
        bs.append(torch.full((d.x.size(0),), g, dtype=torch.long))
        off += d.x.size(0)
    return Data(x=torch.cat(xs), edge_index=torch.cat(eis, 1), y=torch.cat(ys),
                batch=torch.cat(bs), num_graphs=len(graphs))


class GraphDataLoader:
    """``DataLoader(dataset, batch_size, shuffle=True)`` over a list of graphs.  PyG's DataLoader IS
    ``torch.utils.data.DataLoader`` with a graph collate function, so the index batches (and the draws they take from
    the global CPU generator: the iterator's base seed, then the RandomSampler's seed) come from torch's own sampler
    machinery here as well."""

    def __init__(self, dataset, batch_size, shuffle=True):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        import torch.utils.data as tud
        index_batches = tud.DataLoader(range(len(self.dataset)), batch_size=self.batch_size, shuffle=self.shuffle,
                                       collate_fn=list)
        for ids in index_batches:
            yield collate_graphs([self.dataset[i] for i in ids])
