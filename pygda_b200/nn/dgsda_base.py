"""BernProp / DGSDABase -- drop-in for pygda/nn/dgsda_base.py:11-315.

``BernProp.forward`` (:101-153) evaluates  out = sum_k C(K,k)/2^K relu(temp_k) L^k (2I - L)^(K-k) x  with
L = I - D^-1/2 A D^-1/2 (``get_laplacian(..., 'sym')``, :130-131).  The reference recomputes the Laplacian on every
call and spends K + K(K+1)/2 propagations per call (135 at its default K = 15, :135-153).  Here the two graphs
(L and 2I - L) are built once per ``edge_index`` and the same polynomial is evaluated by two Horner sweeps --
T_j = (2I-L)^j x, then R <- L R + a_k T_(K-k) -- i.e. 2K aggregation launches forward and 2K backward (same
kernel as the GCN path, ``gda_spmm_f32``); L and 2I - L commute, so the value is the reference's up to fp32
summation order (tests/test_gpu_dgsda.py: 1e-5 against the reference's own vectors).
"""
import ctypes as C
from math import comb

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .._lib import gda
from ..graph import NORM_SYM_ROW, Graph
from .layers import Linear

_cache = {}


def laplacian_graphs(edge_index, num_nodes):
    """(L, 2I - L) as aggregation graphs for ``edge_index`` -- ``get_laplacian(edge_index, None, 'sym')``
    [upstream PyG: self loops removed, degree at the source index, multi-edges kept] followed by
    ``add_self_loops(edge_index1, -norm1, fill_value=2.)`` (:130-133).  Cached per edge_index tensor."""
    key = getattr(edge_index, "_gda_key", None) or (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape))
    key = (key, int(num_nodes))
    hit = _cache.get(key)
    if hit is not None:
        return hit[0], hit[1]
    ei = edge_index[:, edge_index[0] != edge_index[1]].contiguous()           # remove_self_loops
    ei, w = Graph(ei, num_nodes, None, NORM_SYM_ROW).coo()                      # d^-1/2[row] * 1 * d^-1/2[col]
    loops = torch.arange(num_nodes, dtype=torch.int64, device=ei.device)
    ei_l = torch.cat([ei, loops.unsqueeze(0).repeat(2, 1)], dim=1)
    ones = torch.ones(num_nodes, dtype=torch.float32, device=ei.device)
    lap = Graph(ei_l, num_nodes, torch.cat([-w, ones]), 0)                      # L = I - A_norm
    mid = Graph(ei_l, num_nodes, torch.cat([w, ones]), 0)                       # 2I - L = I + A_norm  (-1 + 2 on the diagonal)
    if len(_cache) >= 16:
        _cache.pop(next(iter(_cache)))
    _cache[key] = (lap, mid, edge_index)                                        # keeps edge_index alive
    return lap, mid


def _p(t):
    return C.c_void_p(t.data_ptr())


class BernPropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, temp, lap, mid, K):
        x = ops._f32c(x)
        temp = ops._f32c(temp)
        coef = [comb(K, k) / (2 ** K) for k in range(K + 1)]
        st = ops._stream
        T = [x]
        for _ in range(K):                                     # T_j = (2I - L)^j x
            T.append(ops.spmm(mid, T[-1]))
        n = x.numel()
        R = torch.empty_like(x)
        gda.bern_axpy_f32(_p(R), _p(T[0]), n, 0.0, coef[K], C.c_void_p(temp.data_ptr() + 4 * K), st())
        for k in range(K - 1, -1, -1):                         # R <- L R + a_k T_(K-k)
            R = ops.spmm(lap, R)
            gda.bern_axpy_f32(_p(R), _p(T[K - k]), n, 1.0, coef[k], C.c_void_p(temp.data_ptr() + 4 * k), st())
        ctx.save_for_backward(temp, *T)
        ctx.cfg = (lap, mid, K, coef)
        return R

    @staticmethod
    def backward(ctx, go):
        temp, *T = ctx.saved_tensors
        lap, mid, K, coef = ctx.cfg
        go = ops._f32c(go)
        st = ops._stream
        n = go.numel()
        G = [go]
        for _ in range(K):                                     # G_k = (L^T)^k g
            G.append(ops.spmm(lap, G[-1], transpose=True))
        gtemp = None
        if ctx.needs_input_grad[1]:
            gtemp = torch.empty(K + 1, dtype=torch.float32, device=go.device)
            scratch = torch.empty(1, dtype=torch.float64, device=go.device)
            for k in range(K + 1):                             # d/d temp_k = c_k [temp_k > 0] <G_k, T_(K-k)>
                gda.bern_dtemp_f32(_p(G[k]), _p(T[K - k]), n, coef[k], C.c_void_p(temp.data_ptr() + 4 * k),
                                   C.c_void_p(gtemp.data_ptr() + 4 * k), _p(scratch), st())
        gx = None
        if ctx.needs_input_grad[0]:
            # dx = sum_k a_k (M^T)^(K-k) G_k: Horner in M^T from the k = 0 term (power K) down to k = K (power 0)
            S = torch.empty_like(go)
            gda.bern_axpy_f32(_p(S), _p(G[0]), n, 0.0, coef[0], C.c_void_p(temp.data_ptr()), st())
            for j in range(K - 1, -1, -1):
                S = ops.spmm(mid, S, transpose=True)
                k = K - j
                gda.bern_axpy_f32(_p(S), _p(G[k]), n, 1.0, coef[k], C.c_void_p(temp.data_ptr() + 4 * k), st())
            gx = S
        return gx, gtemp, None, None, None


class BernProp(nn.Module):
    def __init__(self, K, is_source_domain=True, bias=True, **kwargs):
        super().__init__()
        self.K, self.is_source_domain = K, is_source_domain
        self.cached_terms = None
        self.cached_coefs = None
        self.temp = nn.Parameter(torch.Tensor(self.K + 1), requires_grad=is_source_domain)      # :59
        self.reset_parameters()

    def reset_parameters(self):                                                                 # :63-77
        if self.is_source_domain:
            self.temp.data.fill_(1)
        else:
            self.temp.data = torch.linspace(1, 0, self.K + 1)

    def get_filter(self):                                                                       # :79-99
        TEMP = F.relu(self.temp)
        H = 0
        for k in range(self.K + 1):
            H = H + TEMP[k] * self.cached_coefs[k] * self.cached_terms[k]
        return H

    def forward(self, x, edge_index, edge_weight=None):                                         # :101-153
        if edge_weight is not None:
            raise NotImplementedError("BernProp with edge weights is never used by the reference's DGSDA")
        lap, mid = laplacian_graphs(edge_index, x.size(0))
        return BernPropFn.apply(x, self.temp, lap, mid, self.K)

    def __repr__(self):
        return '{}(K={}, temp={})'.format(self.__class__.__name__, self.K, self.temp)


class DGSDABase(nn.Module):
    """pygda/nn/dgsda_base.py:186-315.  prop1 / prop2 / prop3 are all built with the default
    ``is_source_domain=True`` (:222-224), as in the reference."""

    def __init__(self, features, hidden, classes, dprate=0.0, K=15):
        super().__init__()
        self.lin1 = Linear(features, hidden)
        self.lin2 = Linear(hidden, classes)
        self.prop1 = BernProp(K)
        self.prop2 = BernProp(K)
        self.prop3 = BernProp(K)
        self.dprate = dprate

    def reset_parameters(self):                                                                 # :228-236
        self.prop1.reset_parameters()

    def forward(self, data, is_source_domain=True):                                             # :238-276
        x, edge_index = data.x, data.edge_index
        x = self.get_props(x, edge_index, is_source_domain)
        x = ops.act_dropout(x, None, self.dprate, self.training)
        x = self.lin2(x)
        x = ops.act_dropout(x, None, self.dprate, self.training)
        return self.prop3(x, edge_index)

    def get_props(self, x, edge_index, is_source_domain=True):                                  # :278-315
        x = ops.act_dropout(x, None, self.dprate, self.training)
        x = ops.act_dropout(self.lin1(x), F.relu, self.dprate, self.training)
        return self.prop1(x, edge_index) if is_source_domain else self.prop2(x, edge_index)
