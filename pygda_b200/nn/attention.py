"""Attention -- drop-in for pygda/nn/attention.py:6-55: softmax-weighted fusion of the views' encodings
(used by UDAGCN(ppmi=True), pygda/nn/udagcn_base.py:262-264).  Two views -- the only call site -- run as one
libgda kernel each way (csrc/attention.cu); any other number of views keeps the reference's op sequence."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


class Attention(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.dense_weight = nn.Linear(in_channels, 1)
        self.dropout = nn.Dropout(0.1)                    # constructed and never applied by the reference (:20,52-55)

    def forward(self, inputs):
        if len(inputs) == 2 and inputs[0].is_cuda and inputs[0].dim() == 2:
            return ops.Attention2Fn.apply(inputs[0], inputs[1], self.dense_weight.weight, self.dense_weight.bias)
        # any other number of views: one score per view, softmax across the views, weighted sum (:52-55)
        scores = torch.cat([self.dense_weight(v) for v in inputs], dim=-1)
        alpha = F.softmax(scores, dim=-1)
        out = alpha[..., 0:1] * inputs[0]
        for i in range(1, len(inputs)):
            out = out + alpha[..., i:i + 1] * inputs[i]
        return out
