from oracle.pyg_ops import add_remaining_self_loops, remove_self_loops  # noqa: F401
from oracle.pyg_ops import add_self_loops as _asl
from oracle.pyg_ops import coalesce as _coalesce, dense_to_sparse  # noqa: F401


from oracle.pyg_ops import add_self_loops_attr as _asla, get_laplacian  # noqa: F401,E402


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    if edge_attr is not None and fill_value is not None:      # pygda/nn/dgsda_base.py:133
        return _asla(edge_index, edge_attr, fill_value, num_nodes)
    return _asl(edge_index, num_nodes), edge_attr


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="sum", is_sorted=False, sort_by_row=True):
    # pygda/models/tdss.py:79 calls coalesce(edge_index, None, N, N): the 4th positional lands in `reduce`
    return _coalesce(edge_index, edge_attr, num_nodes)


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    """[1, N, N] with adj[0, row, col] += 1 per edge (pygda/models/strurw.py:509-510)."""
    import torch
    assert batch is None and edge_attr is None
    n = max_num_nodes if max_num_nodes is not None else int(edge_index.max()) + 1
    adj = torch.zeros(n, n)
    adj.index_put_((edge_index[0], edge_index[1]), torch.ones(edge_index.size(1)), accumulate=True)
    return adj.unsqueeze(0)
