"""Adam with torch.optim.Adam's update rule (L2 weight decay, no amsgrad) as ONE
libgda launch over all parameters (gda_adam_step).  Replaces the optimiser the
reference constructs in ``fit`` (pygda/models/a2gnn.py:292-296, 317-319).  Duplicate
parameters in the list receive one update per occurrence, which is what
torch.optim.Adam does with UDAGCN's chained parameter list."""
import ctypes as C

import torch

from ._lib import gda

MAX_TENSORS = 48


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise ValueError("pygda_b200.optim.Adam updates float32 CUDA parameters only")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)

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
        live = [(p, m, v) for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq)
                if p.grad is not None]
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        first = True
        for s in range(0, max(len(live), 1), MAX_TENSORS):
            chunk = live[s:s + MAX_TENSORS]
            n = len(chunk)
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p, _, _ in chunk]
            arr = lambda xs: (C.c_void_p * max(n, 1))(*[x.data_ptr() for x in xs])  # noqa: E731
            numel = (C.c_int64 * max(n, 1))(*[p.numel() for p, _, _ in chunk])
            if not first:
                # the step counter advances once per optimiser step, not per chunk
                self.state[0] -= 1
            gda.adam_step(n, arr([p for p, _, _ in chunk]), arr(grads), arr([m for _, m, _ in chunk]),
                          arr([v for _, _, v in chunk]), numel, float(self.lr), float(self.betas[0]),
                          float(self.betas[1]), float(self.eps), float(self.weight_decay),
                          C.c_void_p(self.state.data_ptr()), stream)
            first = False
