"""Oracle restatement of the un-vendored upstream ops the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  None of the packages below is in
/root/reference; behaviour is restated from torch_geometric 2.4.x /
torch_scatter 2.1 as summarised in SURVEY.md Appendix A, and pinned only by the
dense fp64 tests ("parity unpinned" against the upstream binaries).

Reference call sites each function stands in for are cited per function.
"""
import math

import torch


def scatter_add(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_add(src, index, dim=0, dim_size=N).

    Call sites: pygda/nn/prop_gcn_conv.py:78, pygda/nn/cached_gcn_conv.py:99.
    """
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = (dim_size,) + tuple(src.shape[1:])
    out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    idx = index
    if src.dim() > 1:
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_add_(0, idx, src)


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(edge_index.max()) + 1 if edge_index.numel() else 0


def add_remaining_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    """PyG ``add_remaining_self_loops``.

    Call sites: pygda/nn/prop_gcn_conv.py:72-73, pygda/nn/cached_gcn_conv.py:95-96.
    Non-loop edges keep their order, then loops 0..N-1 are appended; an already
    present loop keeps its own weight.  Multi-edges are not coalesced.
    """
    n = maybe_num_nodes(edge_index, num_nodes)
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    new_index = torch.cat([edge_index[:, keep], loops.unsqueeze(0).repeat(2, 1)], dim=1)
    new_attr = None
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fill_value)
        had = ~keep
        loop_attr[edge_index[0][had]] = edge_attr[had]
        new_attr = torch.cat([edge_attr[keep], loop_attr], dim=0)
    return new_index, new_attr


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, num_nodes):
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loops.unsqueeze(0).repeat(2, 1)], dim=1)


def add_self_loops_attr(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    """PyG ``add_self_loops`` with attributes: N loops are APPENDED (existing loops stay, duplicates are not
    coalesced), their attribute is ``fill_value``.  Call site: pygda/nn/dgsda_base.py:133."""
    n = maybe_num_nodes(edge_index, num_nodes)
    ei = add_self_loops(edge_index, n)
    if edge_attr is None:
        return ei, None
    return ei, torch.cat([edge_attr, edge_attr.new_full((n,), fill_value)], dim=0)


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    """PyG ``get_laplacian`` (upstream, restated from torch_geometric 2.4): self loops removed, degree at the
    SOURCE index, ``'sym'``: L = I - D^-1/2 A D^-1/2 returned as (edges with -w_norm, then N loops with 1);
    ``None``: L = D - A.  Call site: pygda/nn/dgsda_base.py:130-131."""
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    n = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, row, 0, n)
    if normalization is None:
        ei = add_self_loops(edge_index, n)
        return ei, torch.cat([-edge_weight, deg], dim=0)
    if normalization != "sym":
        raise NotImplementedError("only normalization in (None, 'sym') is restated")
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[row] * edge_weight * dis[col]
    return add_self_loops_attr(edge_index, -w, 1.0, n)


def propagate(edge_index, x, edge_weight=None, num_nodes=None):
    """PyG ``MessagePassing.propagate`` with aggr='add', flow source->target
    (SURVEY Appendix A.2): gather rows at edge_index[0], scale, scatter-add at
    edge_index[1].  Call sites: pygda/nn/prop_gcn_conv.py:208-210 (+ message
    :238), pygda/nn/cached_gcn_conv.py:138 (+ message :140-156).
    """
    n = x.size(0) if num_nodes is None else num_nodes
    x_j = x.index_select(0, edge_index[0])
    msg = x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j
    out = torch.zeros((n, x.size(1)), dtype=x.dtype, device=x.device)
    return out.scatter_add_(0, edge_index[1].view(-1, 1).expand_as(msg), msg)


def glorot_(tensor):
    """PyG ``inits.glorot``: U(-a, a), a = sqrt(6 / (size(-2) + size(-1)))."""
    a = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
    with torch.no_grad():
        tensor.uniform_(-a, a)
    return tensor


class Linear(torch.nn.Module):
    """PyG ``Linear(in, out, bias=False, weight_initializer='glorot')``:
    weight [out, in], y = x @ weight.T.  Call site: prop_gcn_conv.py:136-137.
    """

    def __init__(self, in_channels, out_channels, bias=False):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        glorot_(self.weight)

    def forward(self, x):
        y = x @ self.weight.t()
        return y if self.bias is None else y + self.bias


def global_mean_pool(x, batch, size=None):
    """PyG ``global_mean_pool``."""
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    g = int(batch.max()) + 1 if size is None else size
    s = scatter_add(x, batch, 0, g)
    cnt = scatter_add(torch.ones(x.size(0), dtype=x.dtype, device=x.device), batch, 0, g)
    return s / cnt.clamp(min=1).view(-1, 1)


def gcn_norm_by_col(edge_index, edge_weight=None, num_nodes=None, improved=False,
                    add_self_loops_=True, dtype=torch.float32):
    """Tensor branch of pygda/nn/prop_gcn_conv.py:64-81 (== PyG gcn_norm):
    degree accumulated at the TARGET (col) index.
    """
    fill = 2.0 if improved else 1.0
    n = maybe_num_nodes(edge_index, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    if add_self_loops_:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, n)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, col, 0, n)
    dinv = deg.pow(-0.5)
    dinv.masked_fill_(dinv == float("inf"), 0)
    return edge_index, dinv[row] * edge_weight * dinv[col]


def gcn_norm_by_row(edge_index, num_nodes, edge_weight=None, improved=False,
                    dtype=torch.float32):
    """pygda/nn/cached_gcn_conv.py:63-103: as above but the degree is
    accumulated at the SOURCE (row) index (:98-99) and inf is masked (:101).
    """
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    fill = 1.0 if not improved else 2.0
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, row, 0, num_nodes)
    dinv = deg.pow(-0.5)
    dinv[dinv == float("inf")] = 0
    return edge_index, dinv[row] * edge_weight * dinv[col]


def dense_adj(edge_index, edge_weight, num_nodes, dtype=torch.float64):
    """Dense A_hat with A_hat[c, r] = sum of w_e over edges r->c (used only by the
    dense fp64 pin tests)."""
    a = torch.zeros(num_nodes, num_nodes, dtype=dtype)
    a.index_put_((edge_index[1], edge_index[0]), edge_weight.to(dtype), accumulate=True)
    return a


# ---------------------------------------------------------------------------------------------
# Upstream ops behind TDSS's smoothness graph (pygda/models/tdss.py:66-90, 367-383); SURVEY 8(f).1
def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="sum"):
    """PyG ``utils.coalesce`` (2.4): sort by (row, col), drop duplicate pairs (attributes summed).
    Call sites: tdss.py:79, :84 -- called as ``coalesce(edge_index, None, N, N)``, i.e. with
    ``reduce=N`` positionally, which upstream never reads when ``edge_attr`` is None."""
    n = maybe_num_nodes(edge_index, num_nodes)
    key = edge_index[0] * n + edge_index[1]
    key_sorted, perm = torch.sort(key, stable=True)
    uniq, inverse = torch.unique_consecutive(key_sorted, return_inverse=True)
    out_index = torch.stack([uniq // n, uniq % n])
    if edge_attr is None:
        return out_index, None
    attr = edge_attr[perm]
    out_attr = torch.zeros((uniq.numel(),) + tuple(attr.shape[1:]), dtype=attr.dtype)
    out_attr.index_add_(0, inverse, attr)
    return out_index, out_attr


def spspmm(index_a, value_a, index_b, value_b, m, k, n, coalesced=False):
    """``torch_sparse.spspmm``: sparse (m x k) @ sparse (k x n) -> (index [2, nnz] sorted by (row, col),
    value [nnz]).  Call site: tdss.py:73 (A @ A on the unit-weight adjacency; only the PATTERN of the
    result is used -- the values are overwritten with zeros at :74)."""
    import numpy as np
    import scipy.sparse as sp
    a = sp.csr_matrix((value_a.double().numpy(), (index_a[0].numpy(), index_a[1].numpy())), shape=(m, k))
    b = sp.csr_matrix((value_b.double().numpy(), (index_b[0].numpy(), index_b[1].numpy())), shape=(k, n))
    c = (a @ b).tocsr()
    c.sum_duplicates()
    c.sort_indices()
    c = c.tocoo()
    index = torch.from_numpy(np.stack([c.row, c.col]).astype(np.int64))
    return index, torch.from_numpy(c.data).to(value_a.dtype)


def dense_to_sparse(adj):
    """PyG ``dense_to_sparse``: non-zero entries in row-major order.  Call site: tdss.py:373."""
    index = adj.nonzero().t().contiguous()
    return index, adj[index[0], index[1]]


def random_walk(row, col, start, walk_length, generator=None):
    """``torch_cluster.random_walk``: [len(start), walk_length + 1] node ids, each step moving to a
    uniformly chosen out-neighbour (edges row -> col), staying put at nodes without one.
    Call site: tdss.py:370.  The random stream of the upstream C++/CUDA sampler cannot be
    reproduced; parity for the 'RW' smoothing mode is by construction rules only."""
    n = int(max(int(row.max()) if row.numel() else -1, int(col.max()) if col.numel() else -1,
                int(start.max()) if start.numel() else -1)) + 1
    order = torch.argsort(row, stable=True)
    col_sorted = col[order]
    deg = torch.bincount(row, minlength=n)
    ptr = torch.zeros(n + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(deg, 0)
    walk = [start.clone()]
    cur = start.clone()
    for _ in range(walk_length):
        d = deg[cur]
        r = torch.rand(cur.numel(), generator=generator)
        pick = (r * d.clamp(min=1)).long().clamp(max=(d - 1).clamp(min=0))
        nxt = torch.where(d > 0, col_sorted[(ptr[cur] + pick).clamp(max=max(col_sorted.numel() - 1, 0))], cur) \
            if col_sorted.numel() else cur
        walk.append(nxt)
        cur = nxt
    return torch.stack(walk, dim=1)
