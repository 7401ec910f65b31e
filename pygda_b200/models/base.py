"""BaseGDA -- drop-in for pygda/models/base.py:11-163 (same constructor arguments,
``num_neigh`` expansion and errors :86-95, abstract hooks :127-163)."""
from abc import ABC, abstractmethod

import torch.nn.functional as F


class BaseGDA(ABC):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=2, dropout=0., weight_decay=0.,
                 act=F.relu, lr=4e-3, epoch=100, device='cuda:0', batch_size=0, num_neigh=-1,
                 verbose=2, **kwargs):
        super().__init__()
        # hyper-parameters, under the attribute names the reference's estimators and scripts read (base.py:69-84)
        for name, value in (("in_dim", in_dim), ("hid_dim", hid_dim), ("num_classes", num_classes),
                            ("num_layers", num_layers), ("dropout", dropout), ("weight_decay", weight_decay),
                            ("act", act), ("verbose", verbose), ("kwargs", kwargs), ("lr", lr), ("epoch", epoch),
                            ("device", device), ("batch_size", batch_size)):
            setattr(self, name, value)
        self.num_neigh = self._expand_num_neigh(num_neigh, num_layers)
        self.model = None

    @staticmethod
    def _expand_num_neigh(num_neigh, num_layers):
        """One fan-out per layer (base.py:86-95): an int is repeated, a list must have ``num_layers`` entries;
        same exceptions and messages as the reference (bool is rejected there too: ``type(x) is int``)."""
        if type(num_neigh) is int:
            return [num_neigh] * num_layers
        if type(num_neigh) is list:
            if len(num_neigh) == num_layers:
                return num_neigh
            raise ValueError('Number of neighbors should have the '
                             'same length as hidden layers dimension or'
                             'the number of layers.')
        raise ValueError('Number of neighbors must be int or list of int')

    def fit(self, data, **kwargs):
        """Train on the input graph(s)."""

    def predict(self, data, **kwargs):
        """Predict with the fitted model."""

    @abstractmethod
    def init_model(self, **kwargs):
        """Build the torch.nn.Module."""

    @abstractmethod
    def process_graph(self, data, **kwargs):
        """Pre-process the input graph."""

    @abstractmethod
    def forward_model(self, data, **kwargs):
        """One forward pass returning the loss of the batch."""
