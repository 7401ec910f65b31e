"""Attention -- drop-in for pygda/nn/attention.py:6-55 (2-view softmax fusion; only used with
``ppmi=True``).  Tiny elementwise work on [N, 2, H]; kept in torch."""
import torch
import torch.nn.functional as F
from torch import nn


class Attention(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.dense_weight = nn.Linear(in_channels, 1)
        self.dropout = nn.Dropout(0.1)

    def forward(self, inputs):
        stacked = torch.stack(inputs, dim=1)
        weights = F.softmax(self.dense_weight(stacked), dim=1)
        return torch.sum(stacked * weights, dim=1)
