from oracle.pyg_ops import scatter_add  # noqa: F401
