"""Small torch.nn-compatible layers whose math runs on libgda."""
import torch.nn.functional as F
from torch import nn

from .. import ops


class Linear(nn.Linear):
    """``nn.Linear`` (same parameters / init) with the matmul on libgda."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class ReLUDropout(nn.Module):
    """``nn.ReLU()`` followed by ``nn.Dropout(p)`` as one fused kernel; honours train/eval."""

    def __init__(self, p):
        super().__init__()
        self.p = p

    def forward(self, x):
        return ops.act_dropout(x, F.relu, self.p, self.training)


class Dropout(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = p

    def forward(self, x):
        return ops.act_dropout(x, None, self.p, self.training)
