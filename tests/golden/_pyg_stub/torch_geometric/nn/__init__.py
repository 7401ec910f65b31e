from oracle.pyg_ops import global_mean_pool  # noqa: F401
from .conv import MessagePassing, GCNConv  # noqa: F401
