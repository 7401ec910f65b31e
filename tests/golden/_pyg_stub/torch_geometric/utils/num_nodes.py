from oracle.pyg_ops import maybe_num_nodes  # noqa: F401
