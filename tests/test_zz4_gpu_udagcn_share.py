"""UDAGCN(ppmi=True): opt-in sharing of the first-layer product x @ W0 between the adjacency view and the PPMI view
(pygda_b200/nn/udagcn_base.py: ``share_first_product``) -- same encodings and the same gradients as the two separate
evaluations."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


def test_shared_first_product_equals_separate_products():
    from pygda_b200.data import Data
    from pygda_b200.graph import NORM_SYM_ROW, SELF_LOOPS, Graph
    from pygda_b200.nn import UDAGCNBase
    from pygda_b200.synthetic import powerlaw_edge_index
    n, f, h = 3000, 200, 64
    torch.manual_seed(0)
    net = UDAGCNBase(in_dim=f, hid_dim=h, num_classes=4, num_layers=3, ppmi=True).cuda()
    for enc in (net.encoder, net.ppmi_encoder):
        enc.dropout_p = [0.0 for _ in enc.dropout_p]       # the always-on dropout would differ between the two runs
    ei = powerlaw_edge_index(n, 30000, seed=3, offset=2.0).cuda()
    x = torch.randn(n, f).cuda()
    # a fixed weighted graph in place of the (randomly walked) PPMI graph, so that both runs see the same one
    ei2 = powerlaw_edge_index(n, 50000, seed=4, offset=2.0).cuda()
    w2 = (torch.rand(ei2.size(1)) + 0.1).cuda()
    for conv in net.ppmi_encoder.conv_layers:
        conv.cache_dict["c"] = Graph(ei2, n, w2, SELF_LOOPS | NORM_SYM_ROW)
    data = Data(x=x, edge_index=ei)
    go = torch.randn(n, h).cuda()
    params = {id(p): p for m in net.models for p in m.parameters()}
    out = []
    for share in (False, True):
        net.share_first_product = share
        for p in params.values():
            p.grad = None
        enc = net.encode(data, "c")
        enc.backward(go)
        out.append((enc.detach().clone(), {k: p.grad.clone() for k, p in params.items() if p.grad is not None}))
    assert_close(out[1][0], out[0][0], 1e-6, "encoding")
    assert set(out[0][1]) == set(out[1][1])
    for k in out[0][1]:
        assert_close(out[1][1][k], out[0][1][k], 1e-5, "gradient")
    mask = torch.arange(0, n, 3).cuda()
    assert_close(net.encode(data, "c", mask), out[0][0][mask], 1e-6, "masked encoding")
