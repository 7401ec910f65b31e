"""The bf16 feature path (BASELINE.json config 3: "UDAGCN ... hid=256, bf16"): input features and encoder
activations in bf16, parameters / accumulation / losses in fp32.  Tolerances (stated per test): a bf16 rounding is
2^-9 = 2e-3 relative; results that pass through r roundings are compared at ~r x 4e-3 against the fp32 path on
the same inputs, and at fp32 accuracy against fp64 arithmetic on the SAME bf16-rounded operands."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("m,n,k", [(3000, 256, 512), (256, 512, 20000), (1000, 128, 72), (130, 64, 64)])
def test_gemm_bf16_all_majors(ta, tb, m, n, k):
    from pygda_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = _bf(torch.randn((k, m) if ta else (m, k), generator=g)).cuda()
    b = _bf(torch.randn((n, k) if tb else (k, n), generator=g)).cuda()
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    c32 = ops.gemm_bf16(a, b, trans_a=ta, trans_b=tb, out_bf16=False)
    assert c32.dtype == torch.float32
    assert_close(c32, ref, 2e-5, "fp32 output: exact products, fp32 accumulation")
    c16 = ops.gemm_bf16(a, b, trans_a=ta, trans_b=tb, out_bf16=True)
    assert c16.dtype == torch.bfloat16
    assert_close(c16.float(), ref, 4e-3, "bf16 output: one rounding")


def test_small_shapes_take_the_fp32_kernels_on_bf16_values():
    from pygda_b200 import ops
    a, b = _bf(torch.randn(500, 24)).cuda(), _bf(torch.randn(24, 16)).cuda()
    assert_close(ops.gemm_bf16(a, b, out_bf16=False), a.double() @ b.double(), 1e-5, "fallback")


def test_act_dropout_bf16_uses_the_same_masks():
    from pygda_b200 import ops
    x = torch.randn(4001, 256, device="cuda")
    seed, p = 99887766, 0.4
    y32 = ops.ActDropoutFn.apply(x, 1, p, seed)
    xb = _bf(x).requires_grad_(True)
    y16 = ops.ActDropoutFn.apply(xb, 1, p, seed)
    assert y16.dtype == torch.bfloat16
    assert torch.equal(y16 != 0, ops.ActDropoutFn.apply(xb.detach().float(), 1, p, seed) != 0)
    assert_close(y16.float(), y32, 8e-3, "values: two roundings")
    go = _bf(torch.randn(4001, 256, device="cuda"))
    y16.backward(go)
    x32 = xb.detach().float().requires_grad_(True)
    ops.ActDropoutFn.apply(x32, 1, p, seed).backward(go.float())
    assert xb.grad.dtype == torch.bfloat16
    assert_close(xb.grad.float(), x32.grad, 4e-3, "masked gradient")


@pytest.mark.parametrize("w_in_out", [True, False])
def test_graph_conv_bf16_against_the_fp32_node(w_in_out):
    from pygda_b200 import ops
    from pygda_b200.graph import Graph
    from pygda_b200.synthetic import powerlaw_edge_index
    n, fin, h = 6000, 128, 256
    gr = Graph(powerlaw_edge_index(n, 60000, seed=1, offset=2.0).cuda(), n)
    torch.manual_seed(0)
    x = _bf(torch.randn(n, fin)).cuda()
    w = (torch.randn((fin, h) if w_in_out else (h, fin)) * 0.1).cuda().requires_grad_(True)
    b = torch.randn(h).cuda().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = ops.graph_conv(xr, w, b, gr, 1, w_in_out=w_in_out)
    assert y.dtype == torch.bfloat16
    go = _bf(torch.randn(n, h)).cuda()
    y.backward(go)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    x2 = x.float().requires_grad_(True)
    y2 = ops.graph_conv(x2, w2, b2, gr, 1, w_in_out=w_in_out)
    y2.backward(go.float())
    assert_close(y.float(), y2, 1.2e-2, "forward: weights, GEMM output and result rounded to bf16")
    assert_close(w.grad, w2.grad, 1.5e-2, "weight gradient (fp32 accumulation over the nodes)")
    assert_close(b.grad, b2.grad, 1e-3, "bias gradient")
    assert xr.grad.dtype == torch.bfloat16
    assert_close(xr.grad.float(), x2.grad, 1.5e-2, "input gradient")


def test_udagcn_bf16_forward_model_against_the_reference_golden():
    """The reference's own fp32 forward_model vectors (tests/golden/udagcn.pt) at bf16 tolerance."""
    from pygda_b200.data import Data
    from pygda_b200.models import UDAGCN
    g = load_golden("udagcn")
    est = UDAGCN(device="cuda:0", verbose=0, feature_dtype=torch.bfloat16, **g["hparams"])
    est.udagcn = est.init_model()
    est.udagcn.load_state_dict(g["state"])
    est.udagcn.encoder.dropout_p = [0.0 for _ in est.udagcn.encoder.dropout_p]
    est._set_train(False)
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"], g["epoch"])
    loss.backward()
    assert_close(loss, g["loss"], 5e-3, "loss")
    assert_close(s_logits, g["source_logits"], 3e-2, "source logits")
    assert_close(t_logits, g["target_logits"], 3e-2, "target logits")
    for k, p in est.udagcn.named_parameters():
        if k in g["grads"]:
            assert p.grad.dtype == torch.float32
            # gradients are signed sums with cancellation: compared in the Frobenius norm (per-element rounding
            # noise of ~4e-3 per bf16 hop does not shrink with the size of the sum it lands on)
            ref = g["grads"][k].double()
            err = float((p.grad.double().cpu() - ref).norm() / ref.norm().clamp(min=1e-30))
            assert err < 5e-2, f"grad {k}: relative Frobenius error {err:.3e}"


def test_udagcn_bf16_training_tracks_fp32_at_tensor_core_shapes():
    from pygda_b200.models import UDAGCN
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import domain_pair
    import itertools
    src, tgt = domain_pair(6000, 60000, 128, 5, seed=2, device="cuda:0")
    hp = dict(in_dim=128, hid_dim=128, num_classes=5, num_layers=2, ppmi=False, lr=1e-3, weight_decay=1e-3, epoch=100,
              device="cuda:0", verbose=0)
    torch.manual_seed(0)
    a = UDAGCN(**hp)
    a.udagcn = a.init_model()
    b = UDAGCN(feature_dtype=torch.bfloat16, **hp)
    b.udagcn = b.init_model()
    b.udagcn.load_state_dict(a.udagcn.state_dict())
    for est in (a, b):
        est.udagcn.encoder.dropout_p = [0.0 for _ in est.udagcn.encoder.dropout_p]
        est.udagcn.domain_model[1].p = 0.0
    oa = Adam(itertools.chain(*[m.parameters() for m in a.udagcn.models]), lr=1e-3, weight_decay=1e-3)
    ob = Adam(itertools.chain(*[m.parameters() for m in b.udagcn.models]), lr=1e-3, weight_decay=1e-3)
    for step in range(5):
        la, sa, ta, _ = a.train_step(src, tgt, 0.05, step, oa)
        lb, sb, tb, _ = b.train_step(src, tgt, 0.05, step, ob)
        assert_close(lb, la, 1e-2, f"loss step {step}")
        assert_close(sb, sa, 5e-2, f"source logits step {step}")
    assert torch.isfinite(tb).all() and torch.isfinite(lb.detach())
