"""torch_geometric.data.Data stand-in (pygda/models/tdss.py:18,378-379)."""


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kw):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        self.num_nodes = kw.pop("num_nodes", None)
        for k, v in kw.items():
            setattr(self, k, v)
