"""GradReverse -- drop-in for pygda/nn/reverse_layer.py:4-66 (``GradReverse.apply(x, alpha)``)."""
from ..ops import GradReverse  # noqa: F401
