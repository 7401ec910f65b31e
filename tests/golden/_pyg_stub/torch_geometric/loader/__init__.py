from oracle.data import FullBatchNeighborLoader as _FB, GraphDataLoader as DataLoader  # noqa: F401


class NeighborLoader(_FB):
    def __init__(self, data, num_neighbors, batch_size=None, **kw):
        assert batch_size == data.x.shape[0], "stub supports the full-batch mode only"
        super().__init__(data, num_neighbors, batch_size)
