"""Load the reference's REAL hot-path source files from /root/reference without
running ``pygda/__init__.py`` (which imports datasets and every model and so
needs all of PyG).  Only used by make_golden.py, in the build container -- the
GPU box has no /root/reference; tests read the committed fixtures instead.
"""
import importlib
import os
import sys
import types

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def _bare_package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def load_reference():
    """Returns a namespace with the reference's modules for the path."""
    if not os.path.isdir(os.path.join(REF_ROOT, "pygda")):
        raise RuntimeError("reference tree not mounted at /root/reference")
    for p in (REPO, os.path.join(HERE, "_pyg_stub")):
        if p not in sys.path:
            sys.path.insert(0, p)
    base = os.path.join(REF_ROOT, "pygda")
    pkg = _bare_package("pygda", base)
    for sub in ("nn", "utils", "models", "metrics"):
        setattr(pkg, sub, _bare_package("pygda." + sub, os.path.join(base, sub)))

    imp = importlib.import_module
    utility = imp("pygda.utils.utility")
    mmd = imp("pygda.utils.mmd")
    pkg.utils.logger = utility.logger
    pkg.utils.MMD = mmd.MMD
    metrics = imp("pygda.metrics.metrics")
    for k in ("eval_micro_f1", "eval_macro_f1"):
        setattr(pkg.metrics, k, getattr(metrics, k))

    ns = types.SimpleNamespace(mmd=mmd, utility=utility)
    ns.reverse_layer = imp("pygda.nn.reverse_layer")
    pkg.nn.GradReverse = ns.reverse_layer.GradReverse
    ns.prop_gcn_conv = imp("pygda.nn.prop_gcn_conv")
    pkg.nn.PropGCNConv = ns.prop_gcn_conv.PropGCNConv
    ns.a2gnn_base = imp("pygda.nn.a2gnn_base")
    pkg.nn.A2GNNBase = ns.a2gnn_base.A2GNNBase
    ns.cached_gcn_conv = imp("pygda.nn.cached_gcn_conv")
    pkg.nn.CachedGCNConv = ns.cached_gcn_conv.CachedGCNConv
    ns.attention = imp("pygda.nn.attention")

    ns.ppmi_conv = imp("pygda.nn.ppmi_conv")
    pkg.nn.PPMIConv = ns.ppmi_conv.PPMIConv
    ns.udagcn_base = imp("pygda.nn.udagcn_base")
    pkg.nn.UDAGCNBase = ns.udagcn_base.UDAGCNBase
    ns.grade_base = imp("pygda.nn.grade_base")
    pkg.nn.GRADEBase = ns.grade_base.GRADEBase

    ns.adagcn_base = imp("pygda.nn.adagcn_base")
    pkg.nn.AdaGCNBase = ns.adagcn_base.AdaGCNBase

    ns.base = imp("pygda.models.base")
    pkg.models.BaseGDA = ns.base.BaseGDA
    ns.a2gnn = imp("pygda.models.a2gnn")
    ns.udagcn = imp("pygda.models.udagcn")
    ns.grade = imp("pygda.models.grade")
    ns.adagcn = imp("pygda.models.adagcn")
    ns.gnn_base = imp("pygda.nn.gnn_base")
    pkg.nn.GNNBase = ns.gnn_base.GNNBase
    ns.gnn = imp("pygda.models.gnn")
    ns.tdss = imp("pygda.models.tdss")
    ns.dgsda_base = imp("pygda.nn.dgsda_base")
    pkg.nn.DGSDABase = ns.dgsda_base.DGSDABase
    ns.dgsda = imp("pygda.models.dgsda")
    ns.reweight_gnn = imp("pygda.nn.reweight_gnn")
    pkg.nn.ReweightGNN = ns.reweight_gnn.ReweightGNN
    ns.mixup_gcnconv = imp("pygda.nn.mixup_gcnconv")
    pkg.nn.MixUpGCNConv = ns.mixup_gcnconv.MixUpGCNConv
    ns.mixup_base = imp("pygda.nn.mixup_base")
    pkg.nn.MixupBase = ns.mixup_base.MixupBase
    ns.strurw = imp("pygda.models.strurw")
    return ns
