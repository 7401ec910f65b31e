"""DEVELOPER TOOL (see shim_patch.py): __graft_entry__.smoke()'s Python path on the CPU -- the flagship A2GNN step through the
real estimator / module glue with the libgda calls replaced by torch, compared with the oracle."""
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import shim_patch  # noqa
import torch

src = open(os.path.join(ROOT, "__graft_entry__.py")).read().replace("assert torch.cuda.is_available()", "assert True").replace('device="cuda:0"', 'device="cpu"')
ns = {"__file__": os.path.join(ROOT, "__graft_entry__.py"), "__name__": "desk"}
exec(compile(src, "graft_entry_desk", "exec"), ns)
ns["smoke"]()
