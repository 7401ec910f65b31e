"""``fit()`` / ``predict()`` of the estimators on the GPU against the trajectories of the reference's own ``fit``
(tests/golden/fit.pt): same initial weights (injected into the ``init_model`` call inside ``fit``), same CPU RNG state
(MMD indices), the reference's number of epochs -> same final weights and predictions.  Tolerance as in
test_gpu_a2gnn.py::test_train_steps_track_the_oracle: Adam's m / (sqrt(v) + eps) amplifies fp32 noise on entries whose
gradient is close to eps, so weights are compared at 1e-3 of their largest entry."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _inject(est, run):
    real = est.init_model

    def wrapped(**kw):
        net = real(**kw)
        net.load_state_dict(run["init_state"])
        torch.set_rng_state(run["rng_state"])
        return net
    est.init_model = wrapped


@pytest.mark.parametrize("name", ["a2gnn_mmd", "a2gnn_adv"])
def test_a2gnn_fit_reproduces_the_reference_trajectory(name):
    from pygda_b200.data import Data
    from pygda_b200.models import A2GNN
    G = load_golden("fit")
    r = G["runs"][name]
    est = A2GNN(device="cuda:0", verbose=0, **r["hparams"])
    _inject(est, r)
    src, tgt = Data(**G["source"]), Data(**G["target"])              # host-resident, as a user would pass them
    est.fit(src, tgt)
    for k, v in est.a2gnn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-3, "weights after fit: " + k)
    t_logits, t_labels = est.predict(tgt)
    s_logits, s_labels = est.predict(src, source=True)
    assert_close(t_logits, r["target_logits"], 1e-3, "predict(target)")
    assert_close(s_logits, r["source_logits"], 1e-3, "predict(source)")
    assert torch.equal(t_labels.cpu(), r["target_labels"]) and torch.equal(s_labels.cpu(), r["source_labels"])
    # the reference's predict() ignores its ``data`` argument (it re-iterates the loaders fit() stored): so does ours
    assert r["predict_ignores_data"] is True
    other_logits, other_labels = est.predict(src)                     # the SOURCE graph with source=False
    assert torch.equal(other_logits, t_logits) and torch.equal(other_labels, t_labels)


@pytest.mark.parametrize("name", ["strurw_erm", "strurw_adv", "strurw_mmd", "strurw_mixup"])
def test_strurw_fit_reproduces_the_reference_trajectory(name, capsys):
    from pygda_b200.data import Data
    from pygda_b200.models import StruRW
    G = load_golden("fit")
    r = G["runs"][name]
    est = StruRW(device="cuda:0", verbose=0, **r["hparams"])
    _inject(est, r)
    import numpy as np
    np.random.seed(r.get("np_seed", 0))                     # mixup: beta draw + node shuffle per step
    est.fit(Data(**G["source"]), Data(**G["target"]))
    assert capsys.readouterr().out.count("edge reweight...") == 2
    for k, v in est.gnn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-3, "weights after fit: " + k)
    logits, labels = est.predict(Data(**G["target"]))
    assert_close(logits, r["target_logits"], 1e-3, "predict(target)")
    assert torch.equal(labels.cpu(), r["target_labels"])
    if r["hparams"]["mode"] == "adv":
        for k, v in est.domain_discriminator.state_dict().items():
            assert_close(v, r["disc_final_state"][k], 1e-3, "discriminator after fit: " + k)


def _fit_and_compare(est, net_attr, G, r, post=None, source=True, tol=1e-3, resident=False):
    from pygda_b200.data import Data
    real = est.init_model

    def wrapped(**kw):
        net = real(**kw)
        net.load_state_dict(r["init_state"])
        if post is not None:
            post(net)
        torch.set_rng_state(r["rng_state"])
        return net
    est.init_model = wrapped
    src, tgt = Data(**G["source"]), Data(**G["target"])
    if resident:       # GNN.fit never moves its batches (pygda/models/gnn.py:258-262 has no .to): the caller's job
        src, tgt = src.to("cuda:0"), tgt.to("cuda:0")
    est.fit(src, tgt)
    for k, v in getattr(est, net_attr).state_dict().items():
        assert_close(v, r["final_state"][k], tol, "weights after fit: " + k)
    logits, labels = est.predict(tgt)
    assert_close(logits, r["target_logits"], tol, "predict(target)")
    assert torch.equal(labels.cpu(), r["target_labels"])
    if source:
        logits, labels = est.predict(src, source=True)
        assert_close(logits, r["source_logits"], tol, "predict(source)")
        assert torch.equal(labels.cpu(), r["source_labels"])


def test_udagcn_fit_reproduces_the_reference_trajectory():
    from pygda_b200.models import UDAGCN
    G = load_golden("fit")
    r = G["runs"]["udagcn"]

    def no_dropout(net):                                    # as on the reference side (make_golden_fit.py)
        net.encoder.dropout_p = [0.0 for _ in net.encoder.dropout_p]
        net.domain_model[1].p = 0.0
    _fit_and_compare(UDAGCN(device="cuda:0", verbose=0, **r["hparams"]), "udagcn", G, r, post=no_dropout)


@pytest.mark.parametrize("disc", ["js", "mmd"])
def test_grade_fit_reproduces_the_reference_trajectory(disc):
    from pygda_b200.models import GRADE
    G = load_golden("fit")
    r = G["runs"]["grade_" + disc]
    _fit_and_compare(GRADE(device="cuda:0", verbose=0, **r["hparams"]), "grade", G, r)


@pytest.mark.parametrize("backbone", ["gcn", "gat"])
def test_gnn_fit_reproduces_the_reference_trajectory(backbone):
    from pygda_b200.models import GNN
    G = load_golden("fit")
    r = G["runs"]["gnn_" + backbone]
    _fit_and_compare(GNN(device="cuda:0", verbose=0, **r["hparams"]), "gnn", G, r, source=False, resident=True)


def test_tdss_fit_reproduces_the_reference_trajectory():
    from pygda_b200.models import TDSS
    G = load_golden("fit")
    r = G["runs"]["tdss"]
    _fit_and_compare(TDSS(device="cuda:0", verbose=0, **r["hparams"]), "a2gnn", G, r)


def test_dgsda_fit_reproduces_the_reference_trajectory():
    from pygda_b200.models import DGSDA
    G = load_golden("fit")
    r = G["runs"]["dgsda"]
    _fit_and_compare(DGSDA(device="cuda:0", verbose=0, **r["hparams"]), "dgsda", G, r)


@pytest.mark.parametrize("analytic", [False, True])
def test_adagcn_fit_reproduces_the_reference_trajectory(analytic):
    """Three epochs x (10 critic iterations + 1 encoder step); the critic is created inside fit() from the recorded
    generator state, exactly as the reference creates it.  Also with the closed-form critic."""
    from pygda_b200.models import AdaGCN
    G = load_golden("fit")
    r = G["runs"]["adagcn_node"]
    est = AdaGCN(device="cuda:0", verbose=0, **r["hparams"])
    est.analytic_critic = analytic
    real_critic = est.init_critic

    def critic_without_dropout():                          # as on the reference side (make_golden_fit.py)
        real_critic()
        est.discriminator[2].p = 0.0
    est.init_critic = critic_without_dropout

    def encoder_without_dropout(net):                      # the encoder keeps its own Dropout(0.1) (adagcn_base.py:59,84)
        net.encoder.dropout.p = 0.0
    # 33 chained Adam updates (30 of them on the critic): fp32 noise compounds, hence the wider bar
    _fit_and_compare(est, "adagcn", G, r, post=encoder_without_dropout, tol=5e-3)
    for k, v in est.discriminator.state_dict().items():
        assert_close(v, r["critic_final_state"][k], 5e-3, "critic after fit: " + k)


@pytest.mark.parametrize("name", ["a2gnn_graph", "grade_graph"])
def test_graph_mode_minibatch_fit_reproduces_the_reference_trajectory(name):
    """Graph-level mode, DataLoader(batch_size=8, shuffle=True): resident dataset collated on the GPU, batch order from
    torch's sampler on the CPU generator."""
    from pygda_b200.data import Data
    from pygda_b200.models import A2GNN, GRADE
    G = load_golden("fit")
    r = G["runs"][name]
    cls, attr = (A2GNN, "a2gnn") if name.startswith("a2gnn") else (GRADE, "grade")
    est = cls(device="cuda:0", verbose=0, **r["hparams"])
    real = est.init_model

    def wrapped(**kw):
        net = real(**kw)
        net.load_state_dict(r["init_state"])
        torch.set_rng_state(r["rng_state"])
        return net
    est.init_model = wrapped
    est.fit([Data(**d) for d in G["graph_source"]], [Data(**d) for d in G["graph_target"]])
    for k, v in getattr(est, attr).state_dict().items():
        assert_close(v, r["final_state"][k], 1e-3, "weights after fit: " + k)


def test_udagcn_graph_mode_fit_keeps_the_first_batchs_graph_like_the_reference():
    """mode='graph', batch_size=0: one shuffled batch of all graphs per epoch; the conv layers cache the normalised graph
    of the first batch per cache_name and re-use it for the later, differently ordered batches
    (cached_gcn_conv.py:132-136).  The reference's trajectory contains that quirk; so must ours."""
    from pygda_b200.data import Data
    from pygda_b200.models import UDAGCN
    G = load_golden("fit")
    r = G["runs"]["udagcn_graph"]
    est = UDAGCN(device="cuda:0", verbose=0, **r["hparams"])
    real = est.init_model

    def wrapped(**kw):
        net = real(**kw)
        net.load_state_dict(r["init_state"])
        net.encoder.dropout_p = [0.0 for _ in net.encoder.dropout_p]
        net.domain_model[1].p = 0.0
        torch.set_rng_state(r["rng_state"])
        return net
    est.init_model = wrapped
    est.fit([Data(**d) for d in G["graph_source"]], [Data(**d) for d in G["graph_target"]])
    for k, v in est.udagcn.state_dict().items():
        assert_close(v, r["final_state"][k], 1e-3, "weights after fit: " + k)


@pytest.mark.parametrize("name", ["a2gnn_mmd", "gnn_gcn", "strurw_erm"])
def test_fit_prints_the_reference_epoch_lines(name, capsys):
    """verbose=2: 'Epoch NNNN: loss x, source acc y, time z' per epoch -- the summed loss and the micro-F1 of the source
    predictions (training-mode logits for A2GNN, an eval-mode re-prediction after every step for GNN and StruRW), as the
    reference printed them for the same run.  Loss to 1e-3 absolute (printed with four decimals), accuracy exact up to
    one sample of 60."""
    import re
    from pygda_b200.data import Data
    from pygda_b200 import models as M
    G = load_golden("fit")
    r = G["runs"][name]
    cls = {"a2gnn_mmd": M.A2GNN, "gnn_gcn": M.GNN, "strurw_erm": M.StruRW}[name]
    est = cls(device="cuda:0", verbose=2, **r["hparams"])
    _inject(est, r)
    src, tgt = Data(**G["source"]), Data(**G["target"])
    if name == "gnn_gcn":                                  # GNN.fit expects device-resident graphs (gnn.py:258-262)
        src, tgt = src.to("cuda:0"), tgt.to("cuda:0")
    est.fit(src, tgt)
    rows = re.findall(r"Epoch (\d+): loss ([-0-9.]+), source acc ([0-9.]+), time", capsys.readouterr().out)
    got = [(int(e), float(l), float(a)) for e, l, a in rows]
    assert [e for e, _, _ in got] == [e for e, _, _ in r["log"]]
    for (_, loss, acc), (_, rloss, racc) in zip(got, r["log"]):
        assert abs(loss - rloss) <= 1e-3 * max(1.0, abs(rloss)), (got, r["log"])
        assert abs(acc - racc) <= 1.0 / 60 + 1e-4, (got, r["log"])
