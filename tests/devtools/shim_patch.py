"""DEVELOPER TOOL, not a test and not a product path.

When no GPU is at hand (the build container has none, and GPU minutes are rationed), this module lets the PYTHON GLUE
of new estimators / layers be exercised on the CPU: importing it monkeypatches the libgda-backed entry points of
``pygda_b200.ops`` and ``pygda_b200.graph.Graph`` with plain torch expressions inside THIS process only.  It catches
shape / attribute / autograd-wiring mistakes before a GPU call is spent; it proves nothing about the kernels, is never
imported by ``pygda_b200`` or by the pytest suites (``tests/devtools`` holds no ``test_*.py``), and no parity claim
rests on it -- parity is what ``pytest -m gpu`` measures on the B200 against the oracle and the golden vectors.
Usage: ``bash tests/devtools/desk_check.sh tests/test_zz2_gpu_strurw.py``."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from oracle import pyg_ops as P
from oracle import mmd as OM
import pygda_b200
from pygda_b200 import ops, graph as G

SELF_LOOPS, IMPROVED, NORM_SYM_COL, NORM_SYM_ROW = 1, 2, 4, 8

class FakeGraph:
    def __init__(self, edge_index, num_nodes, edge_weight=None, flags=SELF_LOOPS | NORM_SYM_COL):
        self.num_nodes = int(num_nodes)
        ei, w = edge_index, edge_weight
        if flags & NORM_SYM_COL:
            ei, w = P.gcn_norm_by_col(ei, w, num_nodes, bool(flags & IMPROVED), bool(flags & SELF_LOOPS))
        elif (flags & NORM_SYM_ROW) and (flags & SELF_LOOPS):
            ei, w = P.gcn_norm_by_row(ei, num_nodes, w, bool(flags & IMPROVED))
        elif flags & NORM_SYM_ROW:
            assert w is None
            w = torch.ones(ei.size(1))
            deg = P.scatter_add(w, ei[0], 0, num_nodes)
            dinv = deg.pow(-0.5); dinv[dinv == float("inf")] = 0
            w = dinv[ei[0]] * w * dinv[ei[1]]
        elif flags & SELF_LOOPS:
            ei, w = P.add_remaining_self_loops(ei, w if w is not None else torch.ones(ei.size(1)), 1.0, num_nodes)
        elif w is None:
            w = torch.ones(ei.size(1))
        self.ei, self.w = ei, w.float()
        self.nnz = ei.size(1)
    def coo(self):
        return self.ei, self.w
    def csr(self, transpose=False):
        order = torch.argsort(self.ei[1], stable=True)
        return None, self.ei[0][order], self.w[order]

G.Graph = FakeGraph
import pygda_b200.nn.reweight_gnn as RW
RW.Graph = FakeGraph

def spmm(graph, x, transpose=False, bias=None, relu=False, dropout_p=0.0, **kw):
    ei = graph.ei.flip(0) if transpose else graph.ei
    y = P.propagate(ei, x, graph.w, graph.num_nodes)
    if bias is not None: y = y + bias
    if relu: y = torch.relu(y)
    assert dropout_p == 0
    return y
def spmm_k(graph, x, k, transpose=False, bias=None, **kw):
    for i in range(k):
        x = spmm(graph, x, transpose, bias if i == k - 1 else None)
    return x
ops.spmm, ops.spmm_k = spmm, spmm_k

class PropagateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph, k):
        ctx.graph, ctx.k = graph, k
        return spmm_k(graph, x, k)
    @staticmethod
    def backward(ctx, g):
        return spmm_k(ctx.graph, g.contiguous(), ctx.k, transpose=True), None, None
ops.PropagateFn = PropagateFn
RW.ops = ops
def graph_conv(x, weight, bias, graph, k, w_in_out=False):
    h = x @ (weight if w_in_out else weight.t())
    if k > 0:
        h = PropagateFn.apply(h, graph, k)
    return h if bias is None else h + bias
ops.graph_conv = graph_conv
ops.linear = lambda x, w, b=None: F.linear(x, w, b)
ops.matmul = lambda a, b, trans_b=False: a @ (b.t() if trans_b else b)
def act_dropout(x, act, p, training):
    if act is not None: x = act(x)
    return F.dropout(x, p, training) if (training and p > 0) else x
ops.act_dropout = act_dropout
ops.softmax_cross_entropy = lambda z, y: F.cross_entropy(z, y)
def domain_ce(z, split):
    lab = torch.cat([torch.zeros(split, dtype=torch.long), torch.ones(z.size(0) - split, dtype=torch.long)])
    return F.cross_entropy(z, lab)
ops.domain_cross_entropy = domain_ce
ops.combine = lambda pairs: sum(t * w for t, w in pairs)
ops.scale = lambda x, a: x * a
class _BiasAdd:
    @staticmethod
    def apply(x, b):
        return x + b
ops.BiasAddFn = _BiasAdd
import pygda_b200.nn.cached_gcn_conv as _CG
_CG.Graph = FakeGraph
ops.global_mean_pool = lambda x, batch, size=None: P.global_mean_pool(x, batch, size)
import pygda_b200.utils as U
import pygda_b200.models.strurw as S
import pygda_b200.models.adagcn as _A
_A.ops = ops
_A.Adam = torch.optim.Adam
S.MMD = lambda a, b, indices=None: OM.MMD(a, b, indices=indices, sqdist=lambda z: OM.pairwise_sqdist_blocked(z, 250))
S.Adam = torch.optim.Adam
from pygda_b200.metrics import eval_micro_f1
S.micro_f1_from_logits = lambda y, z: eval_micro_f1(y, z.argmax(1))

# LinCombFn: run its real Python (forward/backward) with the two ABI calls emulated on tensors
class _FakeGda:
    @staticmethod
    def scale_f32(y, x, n, a, st):
        y.view(-1)[:n].copy_(a * x.reshape(-1)[:n])
    @staticmethod
    def axpy_f32(y, x, n, a, st):
        y.view(-1)[:n].add_(a * x.reshape(-1)[:n])
ops.gda = _FakeGda
ops._p = lambda t: t
ops._f32c = lambda t: t.contiguous()
ops._stream = lambda: None
ops.scale = lambda x, a: x * a
import pygda_b200.nn.mixup_gcnconv as MG, pygda_b200.nn.mixup_base as MB
MG.ops = ops; MB.ops = ops


def _softmax_entropy(z):
    p = torch.clamp(F.softmax(z, dim=-1), min=1e-9, max=1.0)          # pygda/models/udagcn.py:193-199
    return torch.mean(torch.sum(-p * torch.log(p), dim=-1))
ops.softmax_entropy = _softmax_entropy
import importlib
for _name in ("udagcn", "grade", "a2gnn", "gnn", "dgsda", "tdss"):
    _m = importlib.import_module("pygda_b200.models." + _name)
    if hasattr(_m, "MMD"):
        _m.MMD = lambda a, b, indices=None, **kw: OM.MMD(a, b, indices=indices, sqdist=lambda z: OM.pairwise_sqdist_blocked(z, 250))
    if hasattr(_m, "Adam"):
        _m.Adam = torch.optim.Adam


# ---- A2GNN family: paired bottleneck evaluations and the two-stream overlap, as plain sequential torch ----
def _act_dropout_pair(x, p):
    r = torch.relu(x)
    return F.dropout(r, p, True), F.dropout(r, p, True)


def _graph_conv_act_pair(xa, xb, weight, bias, graph, k, p, relu=True, w_in_out=False):
    def one(x):
        y = graph_conv(x, weight, bias, graph, int(k), w_in_out)
        y = torch.relu(y) if relu else y
        return F.dropout(y, p, True) if p > 0 else y
    return one(xa), one(xb)


ops.act_dropout_pair = _act_dropout_pair
ops.graph_conv_act_pair = _graph_conv_act_pair
import contextlib


class _FakeStream:
    cuda_stream = 0
    def wait_stream(self, other): pass
    def wait_event(self, ev): pass


if not torch.cuda.is_available():
    torch.cuda.current_stream = lambda *a, **k: _FakeStream()
    torch.cuda.Stream = lambda *a, **k: _FakeStream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.Tensor.record_stream = lambda self, s: None
import pygda_b200.optim as _O
_O.Adam = torch.optim.Adam


# ---- TDSS smoothing graph / Laplacian loss, DGSDA's Bernstein propagation: the oracle's forms ----
import pygda_b200.smooth as _SM
from oracle.models import TDSS as _OTDSS
from oracle import nn as _ONN


def _khop(edge_index, num_nodes, k):
    t = _OTDSS(in_dim=1, hid_dim=1, num_classes=2, smooth_mode="K-hop", k=k)
    return t.smoothness(edge_index, None, num_nodes)[0]


_SM.khop_edge_index = _khop
_SM.laplacian_loss = lambda feats, ei: _OTDSS.compute_laplacian_loss(feats, ei)
import pygda_b200.nn.dgsda_base as _DB


def _bern_forward(self, x, edge_index, edge_weight=None):
    o = _ONN.BernProp(self.K)
    o.temp = self.temp
    return o(x, edge_index, edge_weight)


_DB.BernProp.forward = _bern_forward


# the training score: argmax + confusion counts on the GPU in the product, sklearn on the host here
for _name in ("_common", "gnn", "strurw", "a2gnn"):
    _m = importlib.import_module("pygda_b200.models." + _name)
    if hasattr(_m, "micro_f1_from_logits"):
        _m.micro_f1_from_logits = lambda y, z: eval_micro_f1(y, z.argmax(1))
