"""AdaGCN's opt-in closed-form WGAN-GP critic (pygda_b200/models/adagcn.py: ``analytic_critic``): the gradient
penalty without a double backward, the critic's [rows, hid] product on libgda.  Against torch's double backward on
the same inputs, and against the vectors made by the reference's own forward_model (critic weights after its 10
iterations, loss, encoder gradients)."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(400, 400), (300, 500), (500, 300)])
def test_closed_form_penalty_equals_double_backward(shape):
    from pygda_b200.models import AdaGCN
    est = AdaGCN(in_dim=6, hid_dim=128, num_classes=3, adv_dim=40, device="cuda:0", verbose=0)
    torch.manual_seed(0)
    est.init_critic()
    est.discriminator.eval()                              # no dropout: the two forms see the same network
    torch.manual_seed(1)
    es, et = torch.randn(shape[0], 128).cuda(), torch.randn(shape[1], 128).cuda()
    out = []
    for analytic in (False, True):
        est.analytic_critic = analytic
        torch.manual_seed(5)                              # interpolation coefficients (CPU generator)
        gp = est.gradient_penalty(es, et)
        est.discriminator.zero_grad()
        gp.backward()
        out.append((gp.detach().clone(), {k: p.grad.clone() for k, p in est.discriminator.named_parameters()}))
    assert_close(out[1][0], out[0][0], 1e-5, "gradient penalty")
    for k in out[0][1]:
        assert_close(out[1][1][k], out[0][1][k], 2e-4, "grad " + k)
    est.analytic_critic = True
    a = est._critic(es)
    est.analytic_critic = False
    assert_close(a, est.discriminator(es), 1e-5, "critic forward")


@pytest.mark.parametrize("mode", ["node", "graph"])
def test_adagcn_golden_step_with_the_closed_form_critic(mode):
    from pygda_b200.data import Data
    from pygda_b200.models import AdaGCN
    g = load_golden("adagcn_" + mode)
    est = AdaGCN(device="cuda:0", verbose=0, **g["hparams"])
    est.analytic_critic = True
    est.adagcn = est.init_model()
    est.adagcn.load_state_dict(g["state"])
    est.init_critic()
    est.discriminator.load_state_dict(g["critic_state"])
    est.adagcn.eval()
    est.discriminator.eval()
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    torch.manual_seed(g["seed"])
    loss, s_logits, t_logits = est.forward_model(src, tgt)
    est.adagcn.zero_grad()
    loss.backward()
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(s_logits, g["source_logits"], 1e-4, "source logits")
    for k, v in est.discriminator.state_dict().items():
        assert_close(v, g["critic_state_after"][k], 2e-4, "critic after 10 iterations: " + k)
    for k, p in est.adagcn.named_parameters():
        if k in g["grads"]:
            assert_close(p.grad, g["grads"][k], 2e-4, "grad " + k)


def test_node_level_fit_with_the_closed_form_critic():
    from pygda_b200.models import AdaGCN
    from pygda_b200.synthetic import domain_pair
    src, tgt = domain_pair(1500, 10000, 32, 3, seed=5, target_nodes=1300, target_edges=9000)
    torch.manual_seed(0)
    model = AdaGCN(in_dim=32, hid_dim=64, num_classes=3, num_layers=2, dropout=0.1, epoch=2, lr=0.01,
                   device="cuda:0", verbose=0)
    model.analytic_critic = True
    model.fit(src, tgt)
    logits, labels = model.predict(tgt)
    assert logits.shape == (1300, 3) and torch.isfinite(logits).all()
    assert all(torch.isfinite(p).all() for p in model.discriminator.parameters())
