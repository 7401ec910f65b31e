"""pygda_b200 -- B200-native (sm_100a) hot path of PyGDA behind the reference's
``pygda.models`` / ``pygda.nn`` / ``pygda.utils`` API.

    import pygda_b200 as pygda
    model = pygda.models.A2GNN(in_dim=F, hid_dim=128, num_classes=C, device='cuda:0')
    model.fit(source_data, target_data); logits, labels = model.predict(target_data)

All numerics run in hand-written CUDA (pygda_b200/csrc -> libgda.so) through the C ABI
of include/gda.h; there is no CPU fallback.
"""
from . import data, synthetic  # noqa: F401
from . import nn, models, utils, metrics  # noqa: F401
from ._lib import GdaError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
