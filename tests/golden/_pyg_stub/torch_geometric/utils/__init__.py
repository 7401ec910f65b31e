from oracle.pyg_ops import add_remaining_self_loops, remove_self_loops  # noqa: F401
from oracle.pyg_ops import add_self_loops as _asl


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    return _asl(edge_index, num_nodes), edge_attr
