"""``torch_geometric.nn.conv.gcn_conv.gcn_norm`` (pygda/nn/reweight_gnn.py:5,161; pygda/nn/mixup_gcnconv.py:10,204)."""
from oracle import pyg_ops as P


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             flow="source_to_target", dtype=None):
    assert flow == "source_to_target"
    import torch
    return P.gcn_norm_by_col(edge_index, edge_weight, num_nodes, improved, add_self_loops,
                             dtype if dtype is not None else torch.float32)
