"""UDAGCN / GRADE estimators on the GPU vs vectors from the reference's own forward_model."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _check(net, g, loss, s_logits, t_logits):
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    assert_close(t_logits, g["target_logits"], 1e-4, "target logits")
    n = 0
    for k, p in net.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 1e-4, "grad " + k)
            n += 1
    assert n == len(g["grads"])


def test_udagcn_forward_model_golden():
    from pygda_b200.data import Data
    from pygda_b200.models import UDAGCN
    g = load_golden("udagcn")
    est = UDAGCN(device="cuda:0", verbose=0, **g["hparams"])
    est.udagcn = est.init_model()
    est.udagcn.load_state_dict(g["state"])
    est.udagcn.encoder.dropout_p = [0.0 for _ in est.udagcn.encoder.dropout_p]   # see make_golden.py
    est._set_train(False)
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"], g["epoch"])
    _check(est.udagcn, g, loss, s_logits, t_logits)


@pytest.mark.parametrize("disc", ["js", "mmd", "c"])
def test_grade_forward_model_golden(disc):
    from pygda_b200.data import Data
    from pygda_b200.models import GRADE
    g = load_golden("grade_" + disc)
    est = GRADE(device="cuda:0", verbose=0, **g["hparams"])
    est.grade = est.init_model()
    est.grade.load_state_dict(g["state"])
    est.grade.train()
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    torch.manual_seed(g["seed"])              # MMD variant draws its indices from the CPU generator
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"])
    _check(est.grade, g, loss, s_logits, t_logits)


def test_udagcn_fit_predict_and_always_on_encoder_dropout():
    from pygda_b200.models import UDAGCN
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1500, 12000, 48, 3, seed=4)
    torch.manual_seed(0)
    model = UDAGCN(in_dim=48, hid_dim=32, num_classes=3, num_layers=2, ppmi=False, epoch=5, device="cuda:0",
                   verbose=0)
    model.fit(src, tgt)
    a, _ = model.predict(tgt)
    b, _ = model.predict(tgt)
    assert a.shape == (1500, 3)
    assert not torch.equal(a, b)      # the reference's unregistered dropout list stays active in predict


def test_grade_fit_predict_graph_mode():
    from pygda_b200.models import GRADE
    from pygda_b200.synthetic import graph_dataset
    src = graph_dataset(64, 20, 2.0, 14, 2, seed=1)
    tgt = graph_dataset(48, 25, 3.0, 14, 2, seed=2)
    torch.manual_seed(0)
    model = GRADE(in_dim=14, hid_dim=16, num_classes=2, mode="graph", num_layers=2, disc="JS", epoch=3,
                  batch_size=16, device="cuda:0", verbose=0)
    model.fit(src, tgt)
    logits, labels = model.predict(None)
    assert logits.shape == (48, 2) and labels.shape == (48,)


def test_adam_step_counters_follow_each_parameters_own_history():
    """torch.optim.Adam keeps `step` per parameter, creates state at the first gradient and skips parameters without
    one: a parameter that joins late, or misses a step, must get ITS bias corrections."""
    from pygda_b200.optim import Adam
    torch.manual_seed(0)
    shapes = [(7, 5), (5,), (3, 4)]
    ours = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in ours]
    o1 = Adam(ours, lr=0.05, weight_decay=0.01)
    o2 = torch.optim.Adam(ref, lr=0.05, weight_decay=0.01)
    live = [(0, 1), (0, 1, 2), (0, 2), (0, 1, 2), (), (1, 2)]          # which parameters get a gradient at each step
    g = torch.Generator().manual_seed(1)
    for step, idx in enumerate(live):
        for ps in (ours, ref):
            for p in ps:
                p.grad = None
        for i in idx:
            grad = torch.randn(shapes[i], generator=g).cuda()
            ours[i].grad, ref[i].grad = grad.clone(), grad.clone()
        o1.step()
        o2.step()
        for i, (a, b) in enumerate(zip(ours, ref)):
            assert_close(a, b, 1e-5, f"parameter {i} after step {step}")
