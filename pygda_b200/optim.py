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
    def __init__(self, params, mult, dev):
        self.params, self.mult = params, mult
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in params]
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.):
        listed = [p for p in params]
        if not listed:
            raise ValueError("optimizer got an empty parameter list")
        for p in listed:
            if not p.is_cuda or p.dtype != torch.float32:
                raise ValueError("pygda_b200.optim.Adam updates float32 CUDA parameters only")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        dev = listed[0].device
        count, order = {}, []
        for p in listed:
            if id(p) not in count:
                count[id(p)] = 0
                order.append(p)
            count[id(p)] += 1
        self.params = order                                   # unique, in first-occurrence order
        self.groups = [_Group([p for p in order if count[id(p)] == m], m, dev)
                       for m in sorted(set(count.values()))]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                gda.fill_f32(C.c_void_p(p.grad.data_ptr()), p.grad.numel(), 0.0,
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))

    @torch.no_grad()
    def step(self):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for g in self.groups:
            live = [(p, m, v) for p, m, v in zip(g.params, g.exp_avg, g.exp_avg_sq) if p.grad is not None]
            for _ in range(g.mult):
                first = True
                for s in range(0, max(len(live), 1), MAX_TENSORS):
                    chunk = live[s:s + MAX_TENSORS]
                    n = len(chunk)
                    grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p, _, _ in chunk]
                    arr = lambda xs: (C.c_void_p * max(n, 1))(*[x.data_ptr() for x in xs])  # noqa: E731
                    numel = (C.c_int64 * max(n, 1))(*[p.numel() for p, _, _ in chunk])
                    if not first:
                        # the step counter advances once per pass over the group, not per chunk
                        g.state[0] -= 1
                    gda.adam_step(n, arr([p for p, _, _ in chunk]), arr(grads), arr([m for _, m, _ in chunk]),
                                  arr([v for _, _, v in chunk]), numel, float(self.lr), float(self.betas[0]),
                                  float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                  C.c_void_p(g.state.data_ptr()), stream)
                    first = False
