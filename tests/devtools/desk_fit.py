"""DEVELOPER TOOL (see shim_patch.py): fit() / predict() of the main estimators on the CPU (loaders, epoch loop, logging)."""
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import shim_patch  # noqa
import torch
import pygda_b200.models._common as CM
from pygda_b200.metrics import eval_micro_f1
CM.micro_f1_from_logits = lambda y, z: eval_micro_f1(y, z.argmax(1))
from pygda_b200.models import A2GNN, UDAGCN, GRADE, GNN
from pygda_b200.synthetic import domain_pair
src, tgt = domain_pair(600, 4000, 32, 3, seed=2)
torch.manual_seed(0)
for cls, kw in ((A2GNN, dict(s_pnums=0, t_pnums=3, weight=1)), (GRADE, {}), (UDAGCN, dict(ppmi=False))):
    m = cls(in_dim=32, hid_dim=16, num_classes=3, num_layers=2, dropout=0.1, epoch=2, lr=0.01, device="cpu", verbose=2, **kw)
    m.fit(src, tgt)
    logits, labels = m.predict(tgt)
    print(cls.__name__, tuple(logits.shape), bool(torch.isfinite(logits).all()))
