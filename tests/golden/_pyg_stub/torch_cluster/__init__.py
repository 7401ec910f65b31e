"""torch_cluster.random_walk stand-in (pygda/models/tdss.py:15,370): see oracle/pyg_ops.random_walk."""
from oracle.pyg_ops import random_walk  # noqa: F401
