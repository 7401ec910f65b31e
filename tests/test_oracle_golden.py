"""The oracle against the vectors produced by executing the reference's own source
files (tests/golden/make_golden.py).  Index results bit-exact, fp32 within 1e-5."""
import pytest
import torch

from oracle import mmd as OM
from oracle import nn as ONN
from oracle import pyg_ops as P
from oracle.data import Data
from oracle.models import A2GNN as OracleA2GNN
from conftest import assert_close, load_golden


def test_gcn_norm_cases():
    g = load_golden("gcn_norm")
    kw = {"plain": {}, "improved": {"improved": True}, "weighted": {"edge_weight": g["edge_weight"]},
          "noloops": {"add_self_loops_": False}}
    for name, case in g["cases"].items():
        k = kw[name]
        ei, w = P.gcn_norm_by_col(g["edge_index"], k.get("edge_weight"), g["num_nodes"],
                                  k.get("improved", False), k.get("add_self_loops_", True))
        assert torch.equal(ei, case["edge_index_out"]), name
        assert_close(w, case["weight_out"], 1e-6, name)


def test_cached_norm():
    g = load_golden("cached_norm")
    ei, w = P.gcn_norm_by_row(g["edge_index"], g["num_nodes"])
    assert torch.equal(ei, g["edge_index_out"])
    assert_close(w, g["weight_out"], 1e-6, "cached norm")


def test_prop_gcn_conv_forward_backward():
    g = load_golden("prop_gcn_conv")
    conv = ONN.PropGCNConv(12, 8)
    conv.load_state_dict(g["state"])
    for k in (0, 1, 3):
        x = g["x"].clone().requires_grad_(True)
        conv.zero_grad()
        y = conv(x, g["edge_index"], k)
        (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
        assert_close(y, g["out"][k], 1e-5, f"out k={k}")
        assert_close(conv.lin.weight.grad, g["grad_w"][k], 1e-5, f"grad_w k={k}")
        assert_close(x.grad, g["grad_x"][k], 1e-5, f"grad_x k={k}")


def test_cached_gcn_conv():
    g = load_golden("cached_gcn_conv")
    conv = ONN.CachedGCNConv(12, 8)
    conv.load_state_dict(g["state"])
    assert_close(conv(g["x"], g["edge_index"], "c"), g["out"], 1e-5, "cached conv")


def test_mmd_value_grads_and_index_draws():
    g = load_golden("mmd")
    torch.manual_seed(g["seed"])
    s_idx, t_idx = OM.draw_mmd_indices(70, 55, g["sampling_num"], g["times"])
    assert torch.equal(s_idx, g["source_idx"]) and torch.equal(t_idx, g["target_idx"])
    s, t = g["source"].clone().requires_grad_(True), g["target"].clone().requires_grad_(True)
    loss = OM.MMD(s, t, indices=(s_idx, t_idx))
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "mmd loss")
    assert_close(s.grad, g["grad_source"], 1e-5, "mmd grad source")
    assert_close(t.grad, g["grad_target"], 1e-5, "mmd grad target")
    assert_close(OM.get_mmd(g["source"][:50], g["target"][:50]), g["get_mmd_full"], 1e-5, "get_MMD")


def test_grad_reverse():
    g = load_golden("grad_reverse")
    x = g["x"].clone().requires_grad_(True)
    y = ONN.GradReverse.apply(x, g["alpha"])
    y.backward(torch.ones_like(y) * 2.0)
    assert torch.equal(y, g["y"]) and torch.allclose(x.grad, g["grad"])


def _run_a2gnn(name):
    g = load_golden(name)
    est = OracleA2GNN(device="cpu", **g["hparams"])
    est.a2gnn.load_state_dict(g["state"])
    est.a2gnn.train()
    est.mmd_indices = (g["source_idx"], g["target_idx"])
    src, tgt = Data(**g["source"]), Data(**g["target"])
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"])
    est.a2gnn.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, name + " loss")
    assert_close(s_logits, g["source_logits"], 1e-5, name + " source logits")
    assert_close(t_logits, g["target_logits"], 1e-5, name + " target logits")
    for k, p in est.a2gnn.named_parameters():
        assert_close(p.grad, g["grads"][k], 1e-4, name + " grad " + k)


def test_a2gnn_forward_model_mmd():
    _run_a2gnn("a2gnn_mmd")


def test_a2gnn_forward_model_adv():
    _run_a2gnn("a2gnn_adv")


def test_a2gnn_mmd_indices_follow_cpu_generator():
    g = load_golden("a2gnn_mmd")
    torch.manual_seed(g["seed"])
    s_idx, t_idx = OM.draw_mmd_indices(60, 50)
    assert torch.equal(s_idx, g["source_idx"]) and torch.equal(t_idx, g["target_idx"])


def test_logger_lines_equal_the_reference(capsys):
    """pygda_b200.utils.logger prints what the reference's own logger prints (tests/golden/logger.json, made by
    executing pygda/utils/utility.py): the per-epoch line of every fit loop."""
    import json
    import os
    from conftest import GOLDEN
    from pygda_b200.utils import logger
    cases = json.load(open(os.path.join(GOLDEN, "logger.json")))
    assert len(cases) >= 8
    for c in cases:
        kw = dict(c["kwargs"])
        if isinstance(kw.get("loss"), list):
            kw["loss"] = tuple(kw["loss"])
        logger(**kw)
        assert capsys.readouterr().out == c["stdout"], kw


def test_public_signatures_start_with_the_reference_signatures():
    """Drop-in boundary, level 1: every constructor / method of the hot-path classes takes the reference's
    parameters -- same names, order, kinds and defaults (tests/golden/signatures.json, read from the reference's
    own files).  Extra parameters (e.g. ``mmd_indices=None`` for tests) must come last and be optional."""
    import inspect
    import json
    import os
    from conftest import GOLDEN
    import pygda_b200.models as M
    import pygda_b200.nn as NN
    import pygda_b200.utils as U
    ref = json.load(open(os.path.join(GOLDEN, "signatures.json")))
    assert len(ref) >= 90

    def resolve(key):
        parts = key.split(".")
        mod = {"models": M, "nn": NN, "utils": U}[parts[0]]
        obj = getattr(mod, parts[1])
        return getattr(obj, parts[2]) if len(parts) > 2 else obj

    problems = []
    for key, want in ref.items():
        fn = resolve(key)
        got = []
        for p in inspect.signature(fn).parameters.values():
            d = None if p.default is inspect.Parameter.empty else (p.default.__name__ if callable(p.default) else repr(p.default))
            got.append([p.name, p.kind.name, d])
        # a trailing **kwargs of the reference may be followed by nothing; ours may insert optional extras before it
        # (static methods of ours that the reference declares as instance methods differ by `self` only)
        want_core = [w for w in want if w[1] != "VAR_KEYWORD" and w[0] != "self"]
        got_core = [g for g in got if g[1] != "VAR_KEYWORD" and g[0] != "self"]
        if got_core[:len(want_core)] != want_core:
            problems.append((key, want_core, got_core[:len(want_core)]))
            continue
        for extra in got_core[len(want_core):]:
            if extra[2] is None and extra[1] not in ("VAR_POSITIONAL",):
                problems.append((key, "extra parameter without default", extra))
        if any(w[1] == "VAR_KEYWORD" for w in want) and not any(g[1] == "VAR_KEYWORD" for g in got):
            problems.append((key, "reference accepts **kwargs", None))
    assert not problems, problems


def test_estimator_constructors_set_the_reference_attributes():
    """Every simple attribute the reference's constructors set (hyper-parameters, the expanded ``num_neigh`` list,
    mode flags ...) exists with the same value on the pygda_b200 estimator (tests/golden/estimator_attrs.json)."""
    import json
    import os
    from conftest import GOLDEN
    import pygda_b200.models as M
    ref = json.load(open(os.path.join(GOLDEN, "estimator_attrs.json")))
    assert len(ref) == 16
    for key, blob in ref.items():
        est = getattr(M, key.split("/")[0])(**blob["kwargs"])
        for k, v in blob["attrs"].items():
            assert hasattr(est, k), f"{key}: attribute {k} missing"
            assert getattr(est, k) == v, f"{key}: {k} = {getattr(est, k)!r}, reference {v!r}"


def test_num_neigh_validation_errors_like_the_reference():
    import pygda_b200.models as M
    with pytest.raises(ValueError, match="same length"):
        M.A2GNN(in_dim=4, hid_dim=4, num_classes=2, num_layers=2, num_neigh=[1, 2, 3], device="cpu")
    with pytest.raises(ValueError, match="must be int or list"):
        M.GRADE(in_dim=4, hid_dim=4, num_classes=2, num_neigh="all", device="cpu")


def test_module_state_dict_layouts_equal_the_reference():
    """Parameter / buffer names, shapes and trainability of every hot-path module equal the reference's
    (tests/golden/state_dicts.json, from its own files), so reference checkpoints load with strict=True."""
    import json
    import os
    from conftest import GOLDEN
    import pygda_b200.nn as NN
    ref = json.load(open(os.path.join(GOLDEN, "state_dicts.json")))
    assert len(ref) == 21
    for key, blob in ref.items():
        m = getattr(NN, key.split("/")[0])(**blob["kwargs"])
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == blob["state"], key
        assert sorted(k for k, p in m.named_parameters() if p.requires_grad) == blob["trainable"], key


def test_host_metrics_equal_the_reference():
    """pygda_b200.metrics (host-side wrappers) against values computed by the reference's own metrics.py."""
    import json
    import os
    from conftest import GOLDEN
    import pygda_b200.metrics as M
    cases = json.load(open(os.path.join(GOLDEN, "metrics.json")))
    assert len(cases) == 9
    for c in cases:
        label = torch.tensor(c["label"])
        if c["tag"] == "multiclass":
            pred = torch.tensor(c["pred"])
            assert abs(M.eval_micro_f1(label, pred) - c["eval_micro_f1"]) < 1e-12
            assert abs(M.eval_macro_f1(label, pred) - c["eval_macro_f1"]) < 1e-12
            cm = torch.zeros(4, 4, dtype=torch.int64)
            cm.index_put_((label, pred), torch.ones_like(label), accumulate=True)
            assert abs(M.f1_from_confusion(cm, "micro") - c["eval_micro_f1"]) < 1e-12
            assert abs(M.f1_from_confusion(cm, "macro") - c["eval_macro_f1"]) < 1e-12
            continue
        score = torch.tensor(c["score"])
        assert abs(float(M.eval_roc_auc(label, score)) - c["eval_roc_auc"]) < 1e-12
        assert abs(float(M.eval_recall_at_k(label, score)) - c["eval_recall_at_k"]) < 1e-7
        assert abs(float(M.eval_precision_at_k(label, score, 7)) - c["eval_precision_at_k"]) < 1e-7
        assert abs(float(M.eval_average_precision(label, score)) - c["eval_average_precision"]) < 1e-12
