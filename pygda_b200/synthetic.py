"""Seeded synthetic "citation-shaped" graphs.

The reference's datasets are external downloads (pygda/datasets/citation.py,
data/README.md) and unreachable here; the benchmark configs in BASELINE.json are
stated as shapes (nodes / directed edges / features / classes), which these
generators reproduce: undirected, de-duplicated, no self loops, power-law-ish
degree (Chung-Lu endpoint sampling, exponent ~2.5), node ids shuffled; features
non-negative sparse-ish bag-of-words (``relu(randn - shift)``, row-L1
normalised); labels uniform.
"""
import torch

from .data import Data


def powerlaw_edge_index(num_nodes, num_directed_edges, seed=0, gamma=2.5, offset=32.0,
                        device="cpu"):
    """edge_index [2, E] int64 with E == num_directed_edges (even), symmetric."""
    assert num_directed_edges % 2 == 0
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    half = num_directed_edges // 2
    rank = torch.arange(num_nodes, dtype=torch.float64)
    w = (rank + offset).pow(-1.0 / (gamma - 1.0))
    perm = torch.randperm(num_nodes, generator=g)          # hide the degree ordering
    pairs = torch.empty(0, dtype=torch.long)
    while pairs.numel() < half:
        m = int((half - pairs.numel()) * 1.3) + 1024
        u = perm[torch.multinomial(w, m, replacement=True, generator=g)]
        v = perm[torch.multinomial(w, m, replacement=True, generator=g)]
        keep = u != v
        lo, hi = torch.minimum(u, v)[keep], torch.maximum(u, v)[keep]
        key = torch.cat([pairs, lo * num_nodes + hi])
        # unique keeps first occurrence order irrelevant; sort gives determinism
        pairs = torch.unique(key)
        if pairs.numel() > half:
            sel = torch.randperm(pairs.numel(), generator=g)[:half]
            pairs = pairs[sel]
    lo, hi = pairs // num_nodes, pairs % num_nodes
    ei = torch.stack([torch.cat([lo, hi]), torch.cat([hi, lo])])
    order = torch.randperm(ei.size(1), generator=g)        # arbitrary COO order, like raw data
    return ei[:, order].contiguous().to(device)


def powerlaw_edge_index_device(num_nodes, num_directed_edges, seed=0, gamma=2.5, offset=32.0, device="cuda"):
    """The same graph family as ``powerlaw_edge_index`` generated ON the device (inverse-CDF sampling of the Chung-Lu
    endpoint weights): seconds instead of minutes at 5 M nodes / 50 M edges (BASELINE config 4).  Deterministic in
    ``seed`` on a given GPU model, so every rank of a partitioned run builds the same global structure."""
    assert num_directed_edges % 2 == 0
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    half = num_directed_edges // 2
    w = (torch.arange(num_nodes, dtype=torch.float64, device=dev) + offset).pow_(-1.0 / (gamma - 1.0))
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    del w
    perm = torch.randperm(num_nodes, generator=g, device=dev)
    pairs = torch.empty(0, dtype=torch.long, device=dev)
    while pairs.numel() < half:
        m = int((half - pairs.numel()) * 1.3) + 1024
        draw = lambda: perm[torch.searchsorted(cdf, torch.rand(m, generator=g, device=dev, dtype=torch.float64))  # noqa: E731
                            .clamp_(max=num_nodes - 1)]
        u, v = draw(), draw()
        keep = u != v
        lo, hi = torch.minimum(u, v)[keep], torch.maximum(u, v)[keep]
        pairs = torch.unique(torch.cat([pairs, lo * num_nodes + hi]))
        if pairs.numel() > half:
            pairs = pairs[torch.randperm(pairs.numel(), generator=g, device=dev)[:half]]
    lo, hi = pairs // num_nodes, pairs % num_nodes
    ei = torch.stack([torch.cat([lo, hi]), torch.cat([hi, lo])])
    return ei[:, torch.randperm(ei.size(1), generator=g, device=dev)].contiguous()


def bow_features(num_nodes, num_features, seed=0, shift=1.5, device="cpu", chunk=16384):
    """x [N, F] fp32, ``relu(randn - shift)`` row-L1-normalised; generated in row
    chunks on ``device`` so that config-2 size (2.7 GB) never needs a CPU pass."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    x = torch.empty(num_nodes, num_features, dtype=torch.float32, device=dev)
    for s in range(0, num_nodes, chunk):
        e = min(num_nodes, s + chunk)
        blk = torch.randn(e - s, num_features, generator=g, device=dev).sub_(shift).clamp_(min=0)
        blk.div_(blk.sum(1, keepdim=True).clamp_(min=1e-12))
        x[s:e] = blk
    return x


def citation_graph(num_nodes, num_directed_edges, num_features, num_classes, seed=0,
                   feature_shift=1.5, degree_offset=32.0, device="cpu", dtype=torch.float32):
    ei = powerlaw_edge_index(num_nodes, num_directed_edges, seed=seed, offset=degree_offset,
                             device=device)
    x = bow_features(num_nodes, num_features, seed=seed + 1000, shift=feature_shift,
                     device=device).to(dtype)
    g = torch.Generator(device="cpu").manual_seed(int(seed) + 2000)
    y = torch.randint(num_classes, (num_nodes,), generator=g).to(device)
    return Data(x=x, edge_index=ei, y=y)


def domain_pair(num_nodes, num_directed_edges, num_features, num_classes, seed=0,
                device="cpu", target_nodes=None, target_edges=None, dtype=torch.float32):
    """(source, target): same shape family, target with a shifted degree and
    feature distribution (different hub offset / sparsity), different seed."""
    src = citation_graph(num_nodes, num_directed_edges, num_features, num_classes,
                         seed=seed, device=device, dtype=dtype)
    tgt = citation_graph(target_nodes or num_nodes, target_edges or num_directed_edges,
                         num_features, num_classes, seed=seed + 1, feature_shift=1.4,
                         degree_offset=48.0, device=device, dtype=dtype)
    return src, tgt


def graph_dataset(num_graphs, mean_nodes, edges_per_node, num_features, num_classes, seed=0):
    """List of small graphs (Mutagenicity / PROTEINS-shaped, SURVEY section 8d
    config 5): nodes ~ Poisson(mean_nodes) (>= 2), ~edges_per_node directed edges
    per node (symmetric), one-hot features, one label per graph."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    sizes = torch.poisson(torch.full((num_graphs,), float(mean_nodes)), generator=g).long().clamp_(min=2)
    out = []
    for n in sizes.tolist():
        m = max(1, int(round(n * edges_per_node / 2)))
        u = torch.randint(n, (m,), generator=g)
        v = torch.randint(n, (m,), generator=g)
        keep = u != v
        u, v = u[keep], v[keep]
        ei = torch.stack([torch.cat([u, v]), torch.cat([v, u])])
        x = torch.nn.functional.one_hot(torch.randint(num_features, (n,), generator=g),
                                        num_features).float()
        y = torch.randint(num_classes, (1,), generator=g)
        out.append(Data(x=x, edge_index=ei, y=y))
    return out


def community_edge_index(num_blocks, nodes_per_block, edges_per_block, cross_fraction=0.05, seed=0,
                         degree_offset=32.0):
    """Global edge_index of ``num_blocks`` citation-shaped communities (block b owns node ids
    [b*nodes_per_block, (b+1)*nodes_per_block)) with ``edges_per_block`` directed edges each, of
    which ``cross_fraction`` connect to other blocks -- the multi-GPU workload: a graph that
    partitions naturally (one community per GPU, a 5 % edge cut by default).  Deterministic in
    ``seed`` so that every rank can build the same global structure."""
    cross = int(edges_per_block * cross_fraction) // 2 * 2
    intra = edges_per_block - cross
    parts = []
    for b in range(num_blocks):
        ei = powerlaw_edge_index(nodes_per_block, intra, seed=seed + 17 * b, offset=degree_offset)
        parts.append(ei + b * nodes_per_block)
    if num_blocks > 1 and cross > 0:
        g = torch.Generator(device="cpu").manual_seed(int(seed) + 99991)
        m = num_blocks * cross // 2
        ba = torch.randint(num_blocks, (m,), generator=g)
        bb = (ba + 1 + torch.randint(num_blocks - 1, (m,), generator=g)) % num_blocks
        u = ba * nodes_per_block + torch.randint(nodes_per_block, (m,), generator=g)
        v = bb * nodes_per_block + torch.randint(nodes_per_block, (m,), generator=g)
        parts.append(torch.stack([torch.cat([u, v]), torch.cat([v, u])]))
    return torch.cat(parts, 1).contiguous()


def community_block(num_blocks, block, nodes_per_block, edges_per_block, num_features, num_classes, seed=0,
                    cross_fraction=0.05, feature_shift=1.5, degree_offset=32.0, device="cpu"):
    """What one rank of the partitioned run holds: the GLOBAL edge_index and the feature / label
    rows of its own block (as a ``Data`` with ``num_nodes_global``, ``row_lo``, ``row_hi``)."""
    ei = community_edge_index(num_blocks, nodes_per_block, edges_per_block, cross_fraction, seed, degree_offset)
    x = bow_features(nodes_per_block, num_features, seed=seed + 1000 + block, shift=feature_shift, device=device)
    g = torch.Generator(device="cpu").manual_seed(int(seed) + 2000 + block)
    y = torch.randint(num_classes, (nodes_per_block,), generator=g).to(device)
    d = Data(x=x, edge_index=ei.to(device), y=y)
    d.num_nodes_global = num_blocks * nodes_per_block
    d.row_lo, d.row_hi = block * nodes_per_block, (block + 1) * nodes_per_block
    return d
