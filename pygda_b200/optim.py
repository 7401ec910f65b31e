"""Adam with torch.optim.Adam's update rule (L2 weight decay, no amsgrad) as ONE
libgda launch over all parameters (gda_adam_step).  Replaces the optimiser the
reference constructs in ``fit`` (pygda/models/a2gnn.py:292-296, 317-319).

Duplicate parameters: UDAGCN(ppmi=True) chains the parameters of its sub-models into one list in which the
shared conv weights appear twice (pygda/models/udagcn.py:262).  torch.optim.Adam keys
its state by the parameter object, so such a parameter gets TWO sequential updates per step from one shared
state (its step counter advances by two).  Reproduced: parameters are grouped by multiplicity, a group of
multiplicity m is updated by m consecutive launches over the same (exp_avg, exp_avg_sq, step) state."""
import ctypes as C

import torch

from ._lib import gda

MAX_TENSORS = 48


class _Group:
    """Parameters that share one device-side (step count, bias corrections) state: same multiplicity AND the same
    liveness history (they received their first gradient at the same step and have been live together since)."""

    def __init__(self, mult, dev, state=None):
        self.mult = mult
        self.params, self.exp_avg, self.exp_avg_sq = [], [], []
        self.state = torch.zeros(4, dtype=torch.float32, device=dev) if state is None else state

    def add(self, p, m=None, v=None):
        self.params.append(p)
        self.exp_avg.append(torch.zeros_like(p, memory_format=torch.contiguous_format) if m is None else m)
        self.exp_avg_sq.append(torch.zeros_like(p, memory_format=torch.contiguous_format) if v is None else v)

    def pop(self, p):
        i = next(j for j, q in enumerate(self.params) if q is p)
        self.params.pop(i)
        return self.exp_avg.pop(i), self.exp_avg_sq.pop(i)


class Adam:
    """torch.optim.Adam keeps ``step`` per parameter and creates the state lazily on the first step at which
    ``p.grad is not None``; a parameter without a gradient is skipped entirely (its step does not advance).  Here
    parameters are bucketed by that history, so every bucket's shared step counter IS each member's own: a parameter
    joins a bucket at its first gradient (buckets are keyed by multiplicity and birth step), and one that misses a
    step while its bucket-mates are live is split off with a copy of the state.  In the steady state of the fit loops
    (every parameter live on every step) there is one bucket per multiplicity and one launch per bucket."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.):
        listed = [p for p in params]
        if not listed:
            raise ValueError("optimizer got an empty parameter list")
        for p in listed:
            if not p.is_cuda or p.dtype != torch.float32:
                raise ValueError("pygda_b200.optim.Adam updates float32 CUDA parameters only")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.dev = listed[0].device
        count, order = {}, []
        for p in listed:
            if id(p) not in count:
                count[id(p)] = 0
                order.append(p)
            count[id(p)] += 1
        self.params = order                                   # unique, in first-occurrence order
        self.mult = count
        self.groups = []                                      # created lazily, in first-gradient order
        self._of = {}                                         # id(param) -> its bucket
        self._nstep = 0

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                gda.fill_f32(C.c_void_p(p.grad.data_ptr()), p.grad.numel(), 0.0,
                             C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())))

    def _bucket(self):
        """Assign every live parameter to a bucket consistent with its history; returns the buckets to step."""
        fresh = {}
        for p in self.params:
            if p.grad is not None and id(p) not in self._of:
                fresh.setdefault(self.mult[id(p)], []).append(p)
        for m in sorted(fresh):
            g = _Group(m, self.dev)
            self.groups.append(g)
            for p in fresh[m]:
                g.add(p)
                self._of[id(p)] = g
        for g in list(self.groups):
            dead = [p for p in g.params if p.grad is None]
            if dead and len(dead) < len(g.params):            # they fall behind their bucket-mates: split them off
                h = _Group(g.mult, self.dev, g.state.clone())
                for p in dead:
                    h.add(p, *g.pop(p))
                    self._of[id(p)] = h
                self.groups.append(h)
        return [g for g in self.groups if g.params and g.params[0].grad is not None]

    @torch.no_grad()
    def step(self):
        stream = C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
        self._nstep += 1
        for g in self._bucket():
            live = list(zip(g.params, g.exp_avg, g.exp_avg_sq))
            for _ in range(g.mult):
                first = True
                for s in range(0, len(live), MAX_TENSORS):
                    chunk = live[s:s + MAX_TENSORS]
                    n = len(chunk)
                    grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p, _, _ in chunk]
                    arr = lambda xs: (C.c_void_p * n)(*[x.data_ptr() for x in xs])  # noqa: E731
                    numel = (C.c_int64 * n)(*[p.numel() for p, _, _ in chunk])
                    if not first:
                        # the step counter advances once per pass over the group, not per chunk
                        g.state[0] -= 1
                    gda.adam_step(n, arr([p for p, _, _ in chunk]), arr(grads), arr([m for _, m, _ in chunk]),
                                  arr([v for _, _, v in chunk]), numel, float(self.lr), float(self.betas[0]),
                                  float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                  C.c_void_p(g.state.data_ptr()), stream)
                    first = False
