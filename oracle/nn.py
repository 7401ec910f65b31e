"""Oracle restatement of the pygda.nn modules on the hot path (CPU, plain torch).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each class cites the reference
file:line it follows.  State-dict parameter names match the reference so that
weights can be copied between oracle, golden fixtures and the CUDA modules.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import pyg_ops as P


class GradReverse(torch.autograd.Function):
    """pygda/nn/reverse_layer.py:16-66: identity forward (:39), backward
    ``-alpha * g`` (:65-66)."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg() * ctx.alpha, None


class PropGCNConv(nn.Module):
    """pygda/nn/prop_gcn_conv.py:84-264.

    forward (:153-215): gcn_norm recomputed on every call (``cached=False``
    default, :182-192) -> ``lin`` (:205) -> ``prop_nums`` x propagate (:208-210)
    -> in-place ``+= bias`` (:212-213).
    """

    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self._cached_edge_index = None
        self.lin = P.Linear(in_channels, out_channels, bias=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index, prop_nums=1, edge_weight=None):
        if self.normalize:
            cache = self._cached_edge_index
            if cache is None:
                edge_index, edge_weight = P.gcn_norm_by_col(
                    edge_index, edge_weight, x.size(0), self.improved, self.add_self_loops,
                    dtype=x.dtype)
                if self.cached:
                    self._cached_edge_index = (edge_index, edge_weight)
            else:
                edge_index, edge_weight = cache
        out = self.lin(x)
        for _ in range(prop_nums):
            out = P.propagate(edge_index, out, edge_weight)
        if self.bias is not None:
            out = out + self.bias
        return out


class GCNConv(PropGCNConv):
    """Stock PyG ``GCNConv`` = PropGCNConv with exactly one
    propagate.  Call sites: pygda/nn/grade_base.py:58-61, adagcn_base.py:49-52,
    gnn_base.py:65-71."""

    def forward(self, x, edge_index, edge_weight=None):
        return super().forward(x, edge_index, 1, edge_weight)


class A2GNNBase(nn.Module):
    """pygda/nn/a2gnn_base.py:57-203."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, adv=False,
                 dropout=0.1, act=F.relu, mode="node", **kwargs):
        super().__init__()
        self.in_dim, self.hid_dim, self.num_classes = in_dim, hid_dim, num_classes
        self.num_layers, self.adv, self.dropout = num_layers, adv, dropout
        self.act, self.mode = act, mode
        self.convs = nn.ModuleList([PropGCNConv(in_dim, hid_dim)])
        for _ in range(num_layers - 1):
            self.convs.append(PropGCNConv(hid_dim, hid_dim))
        if mode == "node":
            self.cls = PropGCNConv(hid_dim, num_classes)
        else:
            self.cls = nn.Linear(hid_dim, num_classes)
        if adv:
            self.domain_discriminator = nn.Linear(hid_dim, 2)

    def forward(self, data, prop_nums):                       # :72-104
        batch = None if self.mode == "node" else data.batch
        x = self.feat_bottleneck(data.x, data.edge_index, batch, prop_nums=prop_nums)
        return self.feat_classifier(x, data.edge_index, batch, prop_nums=1)

    def feat_bottleneck(self, x, edge_index, batch, prop_nums=30):   # :106-143
        for conv in self.convs:
            x = conv(x, edge_index, prop_nums)
            x = self.act(x)
            x = F.dropout(x, p=self.dropout, training=self.training)
        if self.mode == "graph":
            x = P.global_mean_pool(x, batch)
        return x

    def feat_classifier(self, x, edge_index, batch, prop_nums=1):    # :145-176
        if self.mode == "node":
            return self.cls(x, edge_index, prop_nums)
        return self.cls(x)

    def domain_classifier(self, x, alpha):                            # :178-203
        return self.domain_discriminator(GradReverse.apply(x, alpha))


class CachedGCNConv(nn.Module):
    """pygda/nn/cached_gcn_conv.py:35-174: ``x @ W`` (:130), norm cached per
    ``cache_name`` forever (:132-136, degree by ROW :98-103), one propagate
    (:138), ``+ bias`` in ``update`` (:158-174).  ``weight`` is [in, out]."""

    def __init__(self, in_channels, out_channels, weight=None, bias=None,
                 improved=False, use_bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.improved = in_channels, out_channels, improved
        self.cache_dict = {}
        if weight is None:
            self.weight = nn.Parameter(torch.empty(in_channels, out_channels, dtype=torch.float32))
            P.glorot_(self.weight)
        else:
            self.weight = weight
        if bias is None:
            if use_bias:
                self.bias = nn.Parameter(torch.zeros(out_channels, dtype=torch.float32))
            else:
                self.register_parameter("bias", None)
        else:
            self.bias = bias

    def forward(self, x, edge_index, cache_name="default_cache", edge_weight=None):
        x = torch.matmul(x, self.weight)
        if cache_name not in self.cache_dict:
            edge_index, norm = P.gcn_norm_by_row(edge_index, x.size(0), edge_weight,
                                                 self.improved, x.dtype)
            self.cache_dict[cache_name] = edge_index, norm
        else:
            edge_index, norm = self.cache_dict[cache_name]
        out = P.propagate(edge_index, x, norm)
        if self.bias is not None:
            out = out + self.bias
        return out


class PPMIConv(CachedGCNConv):
    """pygda/nn/ppmi_conv.py:8-184: ``CachedGCNConv`` whose cached graph is the PPMI graph of random walks
    (oracle/ppmi.py restates ``norm``; ``np.random`` is consumed in the reference's order)."""

    def __init__(self, in_channels, out_channels, weight=None, bias=None, improved=False, use_bias=True,
                 path_len=5, **kwargs):
        super().__init__(in_channels, out_channels, weight, bias, improved, use_bias)
        self.path_len = path_len

    def forward(self, x, edge_index, cache_name="default_cache", edge_weight=None):
        from . import ppmi as OP
        x = torch.matmul(x, self.weight)                                   # cached_gcn_conv.py:130
        if cache_name not in self.cache_dict:                              # :132-136 -> PPMIConv.norm
            self.cache_dict[cache_name] = OP.ppmi_norm(edge_index, x.size(0), self.path_len, self.improved)
        edge_index, norm = self.cache_dict[cache_name]
        out = P.propagate(edge_index, x, norm)
        if self.bias is not None:
            out = out + self.bias
        return out


class Attention(nn.Module):
    """pygda/nn/attention.py:6-55."""

    def __init__(self, in_channels):
        super().__init__()
        self.dense_weight = nn.Linear(in_channels, 1)
        self.dropout = nn.Dropout(0.1)

    def forward(self, inputs):
        stacked = torch.stack(inputs, dim=1)
        weights = F.softmax(self.dense_weight(stacked), dim=1)
        return torch.sum(stacked * weights, dim=1)


class UDAGCNEncoder(nn.Module):
    """``GNN`` of pygda/nn/udagcn_base.py:9-89 (gcn type only; the PPMI graph
    builder is out of scope).  ``dropout_layers`` is a plain
    Python list in the reference (:47) so it is never switched to eval mode --
    reproduced here."""

    def __init__(self, in_dim, hid_dim, num_layers=3, base_model=None, act=F.relu, gnn_type="gcn", **kwargs):
        super().__init__()
        cls = PPMIConv if gnn_type == "ppmi" else CachedGCNConv            # udagcn_base.py:51
        if base_model is None:
            weights, biases = [None] * num_layers, [None] * num_layers
        else:
            weights = [c.weight for c in base_model.conv_layers]
            biases = [c.bias for c in base_model.conv_layers]
        self.dropout_layers = [nn.Dropout(0.1) for _ in weights]
        self.act = act
        self.conv_layers = nn.ModuleList()
        self.conv_layers.append(cls(in_dim, hid_dim, weight=weights[0], bias=biases[0], **kwargs))
        for i in range(1, num_layers):
            self.conv_layers.append(cls(hid_dim, hid_dim, weight=weights[i], bias=biases[i], **kwargs))

    def forward(self, x, edge_index, cache_name):
        for i, conv in enumerate(self.conv_layers):
            x = conv(x, edge_index, cache_name)
            if i < len(self.conv_layers) - 1:
                x = self.act(x)
                x = self.dropout_layers[i](x)
        return x


class UDAGCNBase(nn.Module):
    """pygda/nn/udagcn_base.py:92-267 (``ppmi=True``: a second, weight-sharing encoder over PPMIConv layers
    fused with the first by ``Attention``)."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=3, dropout=0.1,
                 act=F.relu, ppmi=False, adv_dim=40, **kwargs):
        super().__init__()
        self.ppmi = ppmi
        self.encoder = UDAGCNEncoder(in_dim, hid_dim, num_layers=num_layers, act=act)
        if ppmi:                                                           # udagcn_base.py:152-153
            self.ppmi_encoder = UDAGCNEncoder(in_dim, hid_dim, num_layers=num_layers, base_model=self.encoder,
                                              gnn_type="ppmi", path_len=10)
        self.cls_model = nn.Sequential(nn.Linear(hid_dim, num_classes))
        self.domain_model = nn.Sequential(
            nn.Linear(hid_dim, adv_dim), nn.ReLU(), nn.Dropout(0.1), nn.Linear(adv_dim, 2))
        self.att_model = Attention(hid_dim)
        self.models = [self.encoder, self.cls_model, self.domain_model]
        if ppmi:                                                           # :168-169
            self.models.extend([self.ppmi_encoder, self.att_model])
        self.loss_func = nn.CrossEntropyLoss()

    def encode(self, data, cache_name, mask=None):                         # :235-267
        out = self.encoder(data.x, data.edge_index, cache_name)
        out = out if mask is None else out[mask]
        if self.ppmi:
            pp = self.ppmi_encoder(data.x, data.edge_index, cache_name)
            pp = pp if mask is None else pp[mask]
            return self.att_model([out, pp])
        return out


class GRADEBase(nn.Module):
    """pygda/nn/grade_base.py:8-202."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, dropout=0.1, act=F.relu,
                 disc="JS", mode="node", **kwargs):
        super().__init__()
        self.num_classes, self.num_layers, self.dropout, self.act, self.mode = \
            num_classes, num_layers, dropout, act, mode
        self.convs = nn.ModuleList([GCNConv(in_dim, hid_dim)])
        for _ in range(num_layers - 1):
            self.convs.append(GCNConv(hid_dim, hid_dim))
        self.cls = nn.Linear(hid_dim, num_classes)
        width = hid_dim * num_layers + (num_classes if disc == "JS" else num_classes * 2)
        self.discriminator = nn.Sequential(nn.Linear(width, 2))
        self.criterion = nn.CrossEntropyLoss()

    def forward(self, data):                                             # :78-115
        batch = None if self.mode == "node" else data.batch
        x, feats = self.feat_bottleneck(data.x, data.edge_index, batch)
        x = self.cls(x)
        feats.append(x)
        return x, torch.cat(feats, dim=1)

    def feat_bottleneck(self, x, edge_index, batch):                     # :117-159
        feats = []
        for conv in self.convs:
            x = conv(x, edge_index)
            x = self.act(x)
            x = F.dropout(x, p=self.dropout, training=self.training)
            feats.append(x if self.mode == "node" else P.global_mean_pool(x, batch))
        if self.mode == "graph":
            x = P.global_mean_pool(x, batch)
        return x, feats

    def one_hot_embedding(self, labels):
        return torch.eye(self.num_classes)[labels]


class AdaGCNEncoder(nn.Module):
    """``GNN`` of pygda/nn/adagcn_base.py:9-59 (gcn type)."""

    def __init__(self, in_dim, hid_dim, num_layers=3, act=F.relu, dropout=0.1):
        super().__init__()
        self.act = act
        self.conv_layers = nn.ModuleList([GCNConv(in_dim, hid_dim)])
        for _ in range(1, num_layers):
            self.conv_layers.append(GCNConv(hid_dim, hid_dim))
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, edge_index, batch, mode="node"):
        for i, conv in enumerate(self.conv_layers):
            x = conv(x, edge_index)
            if i < len(self.conv_layers) - 1:
                x = self.dropout(self.act(x))
        if mode == "graph":
            x = P.global_mean_pool(x, batch)
        return x


class AdaGCNBase(nn.Module):
    """pygda/nn/adagcn_base.py:61-181."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=3, dropout=0.1, act=F.relu,
                 gnn_type="gcn", mode="node", **kwargs):
        super().__init__()
        self.encoder = AdaGCNEncoder(in_dim, hid_dim, num_layers=num_layers, act=act)
        self.cls_model = nn.Sequential(nn.Linear(hid_dim, num_classes))
        self.mode = mode
        self.loss_func = nn.CrossEntropyLoss()

    def forward(self, data):
        batch = None if self.mode == "node" else data.batch
        return self.encoder(data.x, data.edge_index, batch, mode=self.mode)


class GNNBase(nn.Module):
    """gcn branch of pygda/nn/gnn_base.py:8-205."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, dropout=0.1, act=F.relu,
                 gnn="gcn", mode="node", **kwargs):
        super().__init__()
        assert gnn in ("gcn", "gat")
        conv = GCNConv if gnn == "gcn" else (lambda i, o: GATConv(i, o, heads=1, concat=False))
        self.dropout, self.act, self.mode = dropout, act, mode
        self.convs = nn.ModuleList([conv(in_dim, hid_dim)])
        for _ in range(num_layers - 1):
            self.convs.append(conv(hid_dim, hid_dim))
        self.cls = conv(hid_dim, num_classes) if mode == "node" else nn.Linear(hid_dim, num_classes)

    def forward(self, x, edge_index, edge_weight=None, batch=None):
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index, edge_weight)
            if i < len(self.convs) - 1:
                x = F.dropout(self.act(x), p=self.dropout, training=self.training)
        if self.mode == "graph":
            x = P.global_mean_pool(x, batch)
        x = self.cls(x, edge_index, edge_weight) if self.mode == "node" else self.cls(x)
        return F.log_softmax(x, dim=1)


class GATConv(nn.Module):
    """Stock PyG ``GATConv(in, out, heads=1, concat=False)`` (call site
    pygda/nn/gnn_base.py:81-87).  "parity unpinned" upstream restatement: lin_src shared for
    source/target, att_src/att_dst [1,1,out] glorot, negative_slope 0.2, remove + add self loops,
    softmax over the in-edges of each target (max-subtracted, +1e-16), mean over the single head, + bias."""

    def __init__(self, in_channels, out_channels, heads=1, concat=False, negative_slope=0.2, **kwargs):
        super().__init__()
        assert heads == 1 and not concat
        self.out_channels, self.negative_slope = out_channels, negative_slope
        self.lin_src = P.Linear(in_channels, out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        P.glorot_(self.att_src)
        P.glorot_(self.att_dst)

    def forward(self, x, edge_index, edge_weight=None):
        n, c = x.size(0), self.out_channels
        h = self.lin_src(x).view(n, 1, c)
        a_s = (h * self.att_src).sum(-1)
        a_d = (h * self.att_dst).sum(-1)
        ei, _ = P.remove_self_loops(edge_index)
        ei = P.add_self_loops(ei, n)
        e = F.leaky_relu(a_s[ei[0]] + a_d[ei[1]], self.negative_slope)          # [E, 1]
        m = torch.full((n, 1), float("-inf"), dtype=e.dtype).scatter_reduce(0, ei[1].view(-1, 1), e, "amax", include_self=True)
        ex = (e - m[ei[1]]).exp()
        den = P.scatter_add(ex, ei[1], 0, n) + 1e-16
        alpha = ex / den[ei[1]]
        out = P.scatter_add(alpha.unsqueeze(-1) * h[ei[0]], ei[1], 0, n)        # [N, 1, C]
        return out.mean(dim=1) + self.bias


class BernProp(nn.Module):
    """pygda/nn/dgsda_base.py:11-183: Bernstein-polynomial propagation
    ``out = sum_k C(K,k)/2^K relu(temp_k) L^k (2I - L)^(K-k) x`` evaluated the reference's way -- K propagations
    with 2I - L, then k propagations with L for the k-th term (K(K+1)/2 more)."""

    def __init__(self, K, is_source_domain=True, bias=True, **kwargs):
        super().__init__()
        self.K, self.is_source_domain = K, is_source_domain
        self.temp = nn.Parameter(torch.Tensor(K + 1), requires_grad=is_source_domain)   # :59
        self.reset_parameters()

    def reset_parameters(self):                                            # :63-77
        if self.is_source_domain:
            self.temp.data.fill_(1)
        else:
            self.temp.data = torch.linspace(1, 0, self.K + 1)

    def forward(self, x, edge_index, edge_weight=None):                    # :101-153
        from math import comb
        TEMP = F.relu(self.temp)
        ei1, norm1 = P.get_laplacian(edge_index, edge_weight, normalization="sym", dtype=x.dtype,
                                     num_nodes=x.size(0))
        ei2, norm2 = P.add_self_loops_attr(ei1, -norm1, 2.0, x.size(0))
        tmp = [x]
        for _ in range(self.K):
            x = P.propagate(ei2, x, norm2)
            tmp.append(x)
        out = (comb(self.K, 0) / (2 ** self.K)) * TEMP[0] * tmp[self.K]
        for i in range(self.K):
            x = tmp[self.K - i - 1]
            x = P.propagate(ei1, x, norm1)
            for _ in range(i):
                x = P.propagate(ei1, x, norm1)
            out = out + (comb(self.K, i + 1) / (2 ** self.K)) * TEMP[i + 1] * x
        return out


class DGSDABase(nn.Module):
    """pygda/nn/dgsda_base.py:186-315.  NOTE (reference quirk, kept): prop1 / prop2 / prop3 are all built
    with the default ``is_source_domain=True`` (:222-224), so all three ``temp`` vectors start at 1 and train."""

    def __init__(self, features, hidden, classes, dprate=0.0, K=15):
        super().__init__()
        self.lin1 = nn.Linear(features, hidden)
        self.lin2 = nn.Linear(hidden, classes)
        self.prop1, self.prop2, self.prop3 = BernProp(K), BernProp(K), BernProp(K)
        self.dprate = dprate

    def get_props(self, x, edge_index, is_source_domain=True):             # :278-315
        x = F.dropout(x, p=self.dprate, training=self.training)
        x = F.relu(self.lin1(x))
        x = F.dropout(x, p=self.dprate, training=self.training)
        return self.prop1(x, edge_index) if is_source_domain else self.prop2(x, edge_index)

    def forward(self, data, is_source_domain=True):                        # :238-276
        x = self.get_props(data.x, data.edge_index, is_source_domain)
        x = F.dropout(x, p=self.dprate, training=self.training)
        x = self.lin2(x)
        x = F.dropout(x, p=self.dprate, training=self.training)
        return self.prop3(x, data.edge_index)


def _reduce_rows(msg, row, n, aggr):
    """PyG aggregation at ``edge_index[0]`` (flow='target_to_source'): 'add' = scatter-add, 'mean' = the sum
    divided by max(#edges of the row, 1) [upstream MeanAggregation] -- rows without edges give 0."""
    out = P.scatter_add(msg, row, 0, n)
    if aggr == "mean":
        cnt = P.scatter_add(torch.ones(row.numel(), dtype=msg.dtype), row, 0, n)
        out = out / cnt.clamp(min=1).view(-1, 1)
    return out


class GCNReweight(nn.Module):
    """``GCN_reweight`` pygda/nn/reweight_gnn.py:51-249 (ctor :78-113, forward :122-194, message :196-225).
    flow='target_to_source' (:92): gathers x[edge_index[1]], reduces at edge_index[0].  aggr == 'add' switches
    the normalisation OFF (:99-102); otherwise gcn_norm on unit weights without self loops (:161-163; degree at
    edge_index[1], the upstream default flow of gcn_norm).  Message: w x_j ((1 - lmda) + lmda rw) (:223-224)."""

    def __init__(self, in_channels, out_channels, aggr, improved=False, cached=False, add_self_loops=False,
                 normalize=True, bias=True, **kwargs):
        super().__init__()
        self.aggr, self.improved, self.add_self_loops = aggr, improved, add_self_loops
        self.normalize = aggr != "add"
        self.lin = P.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, edge_weight, lmda):
        edge_rw = edge_weight
        edge_weight = torch.ones_like(edge_rw)
        if self.normalize:
            edge_index, edge_weight = P.gcn_norm_by_col(edge_index, edge_weight, x.size(0), self.improved,
                                                        self.add_self_loops, x.dtype)
        x = self.lin(x)
        x_j = edge_weight.view(-1, 1) * x.index_select(0, edge_index[1])
        x_j = (1 - lmda) * x_j + lmda * (edge_rw.view(-1, 1) * x_j)
        out = _reduce_rows(x_j, edge_index[0], x.size(0), self.aggr)
        return out if self.bias is None else out + self.bias


class GSReweight(nn.Module):
    """``GS_reweight`` pygda/nn/reweight_gnn.py:252-372: message = lin(x_j) ((1 - lmda) + lmda w) (:338-339, lin WITH
    bias applied per edge), reduced at edge_index[0] (:275), update = relu(agg_lin(cat(aggr, x))) (+ L2
    normalisation) (:366-372)."""

    def __init__(self, in_channels, out_channels, reducer, normalize_embedding=False):
        super().__init__()
        self.aggr = reducer
        self.lin = nn.Linear(in_channels, out_channels)
        self.agg_lin = nn.Linear(out_channels + in_channels, out_channels)
        self.normalize_emb = normalize_embedding

    def forward(self, x, edge_index, edge_weight, lmda):
        x_j = self.lin(x.index_select(0, edge_index[1]))
        x_j = (1 - lmda) * x_j + lmda * (edge_weight.view(-1, 1) * x_j)
        aggr_out = _reduce_rows(x_j, edge_index[0], x.size(0), self.aggr)
        aggr_out = F.relu(self.agg_lin(torch.cat((aggr_out, x), dim=-1)))
        if self.normalize_emb:
            aggr_out = F.normalize(aggr_out, p=2, dim=-1)
        return aggr_out


class ReweightGNN(nn.Module):
    """pygda/nn/reweight_gnn.py:375-502.  Quirks kept: ONE ``prop_hidden`` module serves every layer after the
    first (:437-446 -- shared weights); ``bns`` are created but never applied (:493-494 commented out);
    ``F.dropout(x, p)`` without ``training=`` (:496) -- dropout stays on in eval mode."""

    def __init__(self, input_dim, gnn_dim, output_dim, cls_dim, gnn_layers=3, cls_layers=2, backbone="GS",
                 pooling="mean", dropout=0.5, bn=False, rw_lmda=1.0, **kwargs):
        super().__init__()
        conv = {"GCN": GCNReweight, "GS": GSReweight}[backbone]
        self.prop_input = conv(input_dim, gnn_dim, pooling)
        self.prop_hidden = conv(gnn_dim, gnn_dim, pooling)
        self.dropout, self.bn, self.lmda = dropout, bn, rw_lmda
        self.conv = nn.ModuleList([self.prop_input] + [self.prop_hidden] * (gnn_layers - 1))
        self.bns = nn.ModuleList([nn.BatchNorm1d(gnn_dim) for _ in range(gnn_layers - 1)])
        self.bn_mlp = nn.BatchNorm1d(cls_dim)
        self.mlp_classify = nn.ModuleList()
        if cls_layers == 1:
            self.mlp_classify.append(nn.Linear(gnn_dim, output_dim))
        else:
            self.mlp_classify.append(nn.Linear(gnn_dim, cls_dim))
            for _ in range(cls_layers - 2):
                self.mlp_classify.append(nn.Linear(cls_dim, cls_dim))
            self.mlp_classify.append(nn.Linear(cls_dim, output_dim))

    def forward(self, data, h):                                            # :462-502
        x, edge_index, edge_weight = h, data.edge_index, data.edge_weight
        for layer in self.conv:
            x = layer(x, edge_index, edge_weight, self.lmda)
            x = F.relu(x)
            x = F.dropout(x, p=self.dropout)
        y = x
        for i, lin in enumerate(self.mlp_classify):
            y = lin(y)
            if i != len(self.mlp_classify) - 1:
                if self.bn:
                    y = self.bn_mlp(y)
                y = F.relu(y)
        return x, y


class MixUpGCNConv(nn.Module):
    """pygda/nn/mixup_gcnconv.py:91-247: out = sum_e w_e ((1 - lmda) + lmda rw_e) lin(x)[row_e] at col_e (default
    flow) + lin_cen(x_cen) + bias, w = gcn_norm WITHOUT self loops on unit weights (:204-206; edge_weight is reset
    to None first, :199-200)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=False,
                 normalize=True, bias=True, **kwargs):
        super().__init__()
        self.improved, self.add_self_loops, self.normalize = improved, add_self_loops, normalize
        self.lin = P.Linear(in_channels, out_channels, bias=False)
        self.lin_cen = P.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, x_cen, edge_index, edge_weight=None, lmda=1):
        edge_rw, edge_weight = edge_weight, None
        if self.normalize:
            edge_index, edge_weight = P.gcn_norm_by_col(edge_index, edge_weight, x.size(0), self.improved,
                                                        self.add_self_loops, x.dtype)
        x = self.lin(x)
        x_j = edge_weight.view(-1, 1) * x.index_select(0, edge_index[0])
        x_j = (1 - lmda) * x_j + lmda * (edge_rw.view(-1, 1) * x_j)
        out = P.scatter_add(x_j, edge_index[1], 0, x.size(0)) + self.lin_cen(x_cen)
        return out if self.bias is None else out + self.bias


class MixupBase(nn.Module):
    """pygda/nn/mixup_base.py:10-200: two-branch mixup over MixUpGCNConv layers (feat_bottleneck :99-178)."""

    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, dropout=0.1, act=F.relu, rw_lmda=0.8, **kwargs):
        super().__init__()
        self.dropout, self.act, self.rw_lmda = dropout, act, rw_lmda
        self.convs = nn.ModuleList([MixUpGCNConv(in_dim, hid_dim)] +
                                   [MixUpGCNConv(hid_dim, hid_dim) for _ in range(num_layers - 1)])
        self.cls = nn.Linear(hid_dim, num_classes)

    def forward(self, x, edge_index, edge_index_b, lam, id_new_value_old, edge_weight):
        return self.cls(self.feat_bottleneck(x, edge_index, edge_index_b, lam, id_new_value_old, edge_weight))

    def feat_classifier(self, x):
        return self.cls(x)

    def feat_bottleneck(self, x, edge_index, edge_index_b, lam, id_new_value_old, edge_weight):
        def drop(t):
            return F.dropout(t, p=self.dropout, training=self.training)

        def conv(i, a, cen, ei):
            return self.convs[i](a, cen, ei, edge_weight, self.rw_lmda)

        perm = torch.as_tensor(id_new_value_old, dtype=torch.long)
        x1 = drop(self.act(conv(0, x, x, edge_index)))                     # :129-135
        x2 = drop(self.act(conv(1, x1, x1, edge_index)))
        x0_b, x1_b = x[perm], x1[perm]                                     # :137-138
        x_mix = x * lam + x0_b * (1 - lam)                                 # :140
        new_x1 = self.act(conv(0, x, x_mix, edge_index))                   # :142-146
        new_x1_b = self.act(conv(0, x0_b, x_mix, edge_index_b))
        x1_mix = drop(new_x1 * lam + new_x1_b * (1 - lam))
        new_x2 = self.act(conv(1, x1, x1_mix, edge_index))                 # :150-156
        new_x2_b = self.act(conv(1, x1_b, x1_mix, edge_index_b))
        x_mix = drop(new_x2 * lam + new_x2_b * (1 - lam))
        x = x2
        for i in range(2, len(self.convs)):                                # :161-176
            x_t = drop(self.act(conv(i, x, x, edge_index)))
            x_b = x[perm]
            new_x = self.act(conv(i, x, x_mix, edge_index))
            new_x_b = self.act(conv(i, x_b, x_mix, edge_index_b))
            x_mix = drop(new_x * lam + new_x_b * (1 - lam))
            x = x_t
        return x_mix
