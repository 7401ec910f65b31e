"""BaseGDA -- drop-in for pygda/models/base.py:11-163 (same constructor arguments,
``num_neigh`` expansion and errors :86-95, abstract hooks :127-163)."""
from abc import ABC, abstractmethod

import torch.nn.functional as F


class BaseGDA(ABC):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=2, dropout=0., weight_decay=0.,
                 act=F.relu, lr=4e-3, epoch=100, device='cuda:0', batch_size=0, num_neigh=-1,
                 verbose=2, **kwargs):
        super().__init__()
        self.in_dim = in_dim
        self.hid_dim = hid_dim
        self.num_classes = num_classes
        self.num_layers = num_layers
        self.dropout = dropout
        self.weight_decay = weight_decay
        self.act = act
        self.verbose = verbose
        self.kwargs = kwargs
        self.lr = lr
        self.epoch = epoch
        self.device = device
        self.batch_size = batch_size
        if type(num_neigh) is int:
            self.num_neigh = [num_neigh] * self.num_layers
        elif type(num_neigh) is list:
            if len(num_neigh) != self.num_layers:
                raise ValueError('Number of neighbors should have the '
                                 'same length as hidden layers dimension or'
                                 'the number of layers.')
            self.num_neigh = num_neigh
        else:
            raise ValueError('Number of neighbors must be int or list of int')
        self.model = None

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
