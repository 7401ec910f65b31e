"""CPU oracle for the PyGDA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in plain PyTorch on the CPU, the op sequence the
reference (pygda-team/pygda @ 8cef1de, mounted read-only at /root/reference)
executes on its per-training-step hot path: GCN normalisation, COO
gather/scatter aggregation, the A2GNN / UDAGCN / GRADE / AdaGCN encoders, the
Gaussian-MMD loss and the gradient-reversal discriminators.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- only as the
checker or the CPU baseline.  Nothing under ``pygda_b200/`` imports it; the
product path fails loudly when its CUDA library is missing.

Parity pinning status
---------------------
The reference ships no tests, golden vectors or fixtures
and cannot be imported as a package here because torch_geometric /
torch_scatter / torch_sparse are not installed.  Pinning is therefore two-tier:

* PINNED against the reference's own code: ``tests/golden/make_golden.py``
  executes the reference's real source files for the path
  (``pygda/utils/mmd.py``, ``pygda/nn/reverse_layer.py``,
  ``pygda/nn/prop_gcn_conv.py``, ``pygda/nn/a2gnn_base.py``, ...) from
  /root/reference with a minimal stand-in for the three missing third-party
  packages, and commits the input/output vectors under ``tests/golden/``.
  ``tests/test_oracle_golden.py`` checks this oracle against those vectors.
* UNPINNED ("parity unpinned" for these): the semantics of the un-vendored
  upstream ops themselves (PyG ``MessagePassing.propagate``,
  ``add_remaining_self_loops``, ``Linear``/glorot, ``global_mean_pool``,
  ``GATConv``; ``torch_scatter.scatter_add``) are restated from their published
  2.4.x behaviour and checked only against dense fp64
  linear algebra (``tests/test_oracle_dense.py``), not against the upstream
  binaries.
"""

from . import pyg_ops, nn, mmd, data, models  # noqa: F401
