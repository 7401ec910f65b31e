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
