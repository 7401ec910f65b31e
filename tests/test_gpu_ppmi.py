"""PPMI graph construction on the GPU (pygda_b200/ppmi.py, csrc/ppmi.cu) against the oracle's restatement of
PPMIConv.norm (oracle/ppmi.py, pinned by the reference's own file through tests/golden/ppmi.pt).

NumPy's random stream cannot be reproduced on the GPU, so parity is checked in two halves: (1) the walks are valid
samples of the reference's procedure (uniform neighbour steps on the de-duplicated undirected graph, uniform
length in [1, path_len], 40 rounds from every node that has an edge) and the exported visit counts are exactly
those of the walks; (2) given the visit counts, scores / normalisation / the UDAGCN(ppmi=True) forward are the
reference's arithmetic to fp32 round-off."""
from collections import Counter

import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import ppmi as OP

pytestmark = pytest.mark.gpu


def _graph(n=300, e=1500, seed=0):
    from pygda_b200.synthetic import powerlaw_edge_index
    ei = powerlaw_edge_index(n, e, seed=seed, offset=2.0)
    ei = ei[:, : e // 2]                                     # directed input: the builder symmetrises it
    ei = torch.cat([ei, torch.tensor([[5, 5, 7], [5, 5, 9]])], 1)          # a self loop, duplicates
    ei = ei[:, (ei[0] < n - 3) & (ei[1] < n - 3)]            # the last three nodes have no edge at all
    return ei, n


def test_walks_are_valid_and_counts_are_theirs():
    from pygda_b200.ppmi import ppmi_edges, ppmi_walks
    ei, n = _graph()
    L, rounds, seed = 6, 40, 1234
    walks = ppmi_walks(ei.cuda(), n, L, rounds, seed=seed).cpu()
    edges, w, cnt = ppmi_edges(ei.cuda(), n, L, rounds, seed=seed, return_counts=True)
    edges, cnt = edges.cpu(), cnt.cpu()
    adj = {}
    for a, b in ei.t().tolist():
        adj.setdefault(a, set()).add(b)
        adj.setdefault(b, set()).add(a)
    ref = Counter()
    lengths = []
    for r in range(rounds):
        for a in range(n):
            row = walks[r, a].tolist()
            if a not in adj:
                assert all(v == -1 for v in row)             # `for a in adj_dict`: no walk
                continue
            k = sum(v >= 0 for v in row)
            assert 1 <= k <= L and all(v >= 0 for v in row[:k]) and all(v == -1 for v in row[k:])
            lengths.append(k)
            cur = a
            for v in row[:k]:
                assert v in adj[cur], "a step must follow an (undirected) edge"
                ref[(a, v)] += 1
                cur = v
    # (2) counts exported by the builder == counts of these very walks; edges sorted by (start, visited)
    got = {(int(a), int(b)): int(c) for a, b, c in zip(edges[0], edges[1], cnt)}
    assert got == dict(ref)
    key = edges[0] * n + edges[1]
    assert bool((key[1:] > key[:-1]).all())
    # walk length ~ uniform{1..L}
    hist = np.bincount(lengths, minlength=L + 1)[1:] / len(lengths)
    assert np.abs(hist - 1.0 / L).max() < 0.02
    # neighbour choice ~ uniform: first steps out of the best-connected node
    hub = max(adj, key=lambda a: len(adj[a]))
    big = ppmi_walks(ei.cuda(), n, 1, 4000, seed=99)[:, hub, 0].cpu().tolist()
    freq = Counter(big)
    assert set(freq) <= adj[hub]
    exp = 4000 / len(adj[hub])
    chi2 = sum((freq.get(v, 0) - exp) ** 2 / exp for v in adj[hub])
    assert chi2 < 2.0 * len(adj[hub]) + 20                   # mean of chi2 = deg - 1
    # reproducible for a seed, different for another
    again = ppmi_walks(ei.cuda(), n, L, rounds, seed=seed).cpu()
    other = ppmi_walks(ei.cuda(), n, L, rounds, seed=seed + 1).cpu()
    assert torch.equal(again, walks) and not torch.equal(other, walks)


def test_scores_and_normalisation_match_the_oracle_given_the_counts():
    from pygda_b200.nn import PPMIConv
    from pygda_b200.ppmi import ppmi_edges
    ei, n = _graph(400, 3000, seed=2)
    L = 10
    edges, w, cnt = ppmi_edges(ei.cuda(), n, L, seed=7, return_counts=True)
    ref_w = OP.ppmi_from_counts(edges[0].cpu().numpy(), edges[1].cpu().numpy(), cnt.cpu().numpy(), L)
    assert (ref_w == 0).any() and (ref_w > 0).any()          # zero scores are kept as edges (ppmi_conv.py:165-170)
    assert_close(w, torch.from_numpy(ref_w), 1e-6, "ppmi scores")
    # the whole norm(): remaining self loops (an existing loop keeps ITS score, even 0), D^-1/2 W D^-1/2 by row
    conv = PPMIConv(8, 4, path_len=L)
    torch.manual_seed(3)
    g = conv._ppmi_graph(ei.cuda(), n)
    got_ei, got_w = g.coo()
    torch.manual_seed(3)
    e2, w2, c2 = ppmi_edges(ei.cuda(), n, L, return_counts=True)           # same CPU-generator seed -> same walks
    w64 = torch.from_numpy(OP.ppmi_from_counts(e2[0].cpu().numpy(), e2[1].cpu().numpy(), c2.cpu().numpy(), L))
    ref_ei, ref_norm = OP.sym_norm_by_row(e2.cpu(), w64, n)
    assert torch.equal(got_ei.cpu(), ref_ei)
    assert_close(got_w, ref_norm, 1e-5, "normalised ppmi weights")


def test_visit_distribution_matches_the_reference_procedure():
    """Same procedure, different random stream: per-start visit probabilities of many GPU rounds against many
    rounds of the oracle's line-by-line restatement (np.random)."""
    from pygda_b200.ppmi import ppmi_edges
    ei, n = _graph(40, 160, seed=5)
    L = 5
    edges, _, cnt = ppmi_edges(ei.cuda(), n, L, rounds=4000, seed=11, return_counts=True)
    gp = torch.zeros(n, n, dtype=torch.float64)
    gp[edges[0].cpu(), edges[1].cpu()] = cnt.cpu().double()
    np.random.seed(0)
    counters = OP.walk_counters(ei, L, rounds=4000)
    rp = torch.zeros(n, n, dtype=torch.float64)
    for a, c in counters.items():
        for b, k in c.items():
            rp[a, b] = k
    assert torch.equal(gp.sum(1) > 0, rp.sum(1) > 0)
    assert abs(float(gp.sum()) / float(rp.sum()) - 1) < 0.02              # total steps: rounds x starts x (L+1)/2
    act = gp.sum(1) > 0
    gpn, rpn = gp[act] / gp[act].sum(1, keepdim=True), rp[act] / rp[act].sum(1, keepdim=True)
    assert float((gpn - rpn).abs().max()) < 0.03


def test_udagcn_ppmi_forward_matches_the_oracle_on_the_same_ppmi_graphs():
    from oracle.data import Data as OData
    from oracle.models import UDAGCN as OUDAGCN
    from pygda_b200.data import Data
    from pygda_b200.models import UDAGCN
    g = load_golden("ppmi")["udagcn_ppmi"]
    est = UDAGCN(device="cuda:0", verbose=0, **g["hparams"])
    net = est.udagcn = est.init_model()
    net.load_state_dict(g["state"])
    net.encoder.dropout_p = [0.0 for _ in net.encoder.dropout_p]
    net.ppmi_encoder.dropout_p = [0.0 for _ in net.ppmi_encoder.dropout_p]
    est._set_train(False)
    src, tgt = Data(**g["source"]).to("cuda:0"), Data(**g["target"]).to("cuda:0")
    torch.manual_seed(0)
    loss, s_logits, t_logits = est.forward_model(src, tgt, g["alpha"], g["epoch"])
    loss.backward()

    ora = OUDAGCN(**g["hparams"])
    onet = ora.udagcn
    onet.load_state_dict(g["state"])
    onet.encoder.dropout_layers = [torch.nn.Identity() for _ in onet.encoder.dropout_layers]
    onet.ppmi_encoder.dropout_layers = [torch.nn.Identity() for _ in onet.ppmi_encoder.dropout_layers]
    for m in onet.models:
        m.eval()
    for li, conv in enumerate(net.ppmi_encoder.conv_layers):               # inject the GPU-built PPMI graphs
        assert set(conv.cache_dict) == {"source", "target"}
        for name, graph in conv.cache_dict.items():
            ei, w = graph.coo()
            onet.ppmi_encoder.conv_layers[li].cache_dict[name] = (ei.cpu(), w.cpu())
    rl, rs, rt = ora.forward_model(OData(**g["source"]), OData(**g["target"]), g["alpha"], g["epoch"])
    onet.zero_grad()
    rl.backward()
    assert_close(loss, rl, 1e-4, "loss")
    assert_close(s_logits, rs, 1e-4, "source logits")
    assert_close(t_logits, rt, 1e-4, "target logits")
    ograds = {k: p.grad for k, p in onet.named_parameters() if p.grad is not None}
    n = 0
    for k, p in net.named_parameters():
        if k in ograds:
            if k == "att_model.dense_weight.bias":        # both views share it: softmax is shift invariant, grad == 0
                assert float(p.grad.abs().max()) < 1e-6 and float(ograds[k].abs().max()) < 1e-6
            else:
                assert_close(p.grad, ograds[k], 1e-4, "grad " + k)
            n += 1
    assert n == len(ograds) and n >= 10
    # and it is in the same regime as the reference's own run (different walks): loss within a few per cent
    assert abs(float(loss) - float(g["loss"])) < 0.1 * abs(float(g["loss"]))


def test_two_view_attention_kernel_matches_the_op_sequence():
    """csrc/attention.cu against stack -> Linear -> softmax -> weighted sum (pygda/nn/attention.py:52-55)."""
    import torch.nn.functional as F
    from pygda_b200.nn import Attention
    torch.manual_seed(0)
    for n, h in ((1000, 128), (77, 16), (513, 200)):
        att = Attention(h).cuda()
        x0 = torch.randn(n, h, device="cuda", requires_grad=True)
        x1 = torch.randn(n, h, device="cuda", requires_grad=True)
        out = att([x0, x1])
        go = torch.randn_like(out)
        out.backward(go)
        got = [x0.grad.clone(), x1.grad.clone(), att.dense_weight.weight.grad.clone(), att.dense_weight.bias.grad.clone()]
        y0, y1 = x0.detach().double().requires_grad_(True), x1.detach().double().requires_grad_(True)
        w = att.dense_weight.weight.detach().double().requires_grad_(True)
        b = att.dense_weight.bias.detach().double().requires_grad_(True)
        stacked = torch.stack([y0, y1], dim=1)
        ref = torch.sum(stacked * F.softmax(stacked @ w.t() + b, dim=1), dim=1)
        ref.backward(go.double())
        assert_close(out, ref, 1e-5, "attention forward")
        for name, g, r in zip(("x0", "x1", "weight", "bias"), got, (y0.grad, y1.grad, w.grad, b.grad)):
            if name == "bias":
                assert float(g.abs().max()) == 0.0 and float(r.abs().max()) < 1e-9
            else:
                assert_close(g, r, 1e-4, "attention grad " + name)


def test_udagcn_default_fit_and_duplicate_parameter_updates():
    """ppmi=True is the reference's default; its chained parameter list holds the shared conv weights twice and
    torch's Adam then applies two updates per step from one state -- reproduced by
    pygda_b200.optim.Adam: compare with torch.optim.Adam on the same duplicated list."""
    from pygda_b200.models import UDAGCN
    from pygda_b200.optim import Adam
    from pygda_b200.synthetic import domain_pair
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(64, 8, device="cuda"))
    b = torch.nn.Parameter(torch.randn(5, device="cuda"))
    a2, b2 = torch.nn.Parameter(a.detach().clone()), torch.nn.Parameter(b.detach().clone())
    mine = Adam([a, b, a], lr=0.01, weight_decay=0.003)
    # foreach=False: the sequential per-parameter loop (torch's CPU default, what the oracle runs and what the
    # survey verified); the CUDA multi-tensor path races on a list that holds one tensor twice
    with pytest.warns(UserWarning):
        ref = torch.optim.Adam([a2, b2, a2], lr=0.01, weight_decay=0.003, foreach=False)
    for step in range(3):
        ga, gb = torch.randn_like(a), torch.randn_like(b)
        a.grad, b.grad, a2.grad, b2.grad = ga.clone(), gb.clone(), ga.clone(), gb.clone()
        mine.step()
        ref.step()
        assert_close(a, a2, 1e-5, f"duplicated parameter, step {step}")
        assert_close(b, b2, 1e-5, f"plain parameter, step {step}")
    src, tgt = domain_pair(1200, 9000, 48, 3, seed=4)
    torch.manual_seed(0)
    model = UDAGCN(in_dim=48, hid_dim=32, num_classes=3, num_layers=2, epoch=4, device="cuda:0", verbose=0)
    assert model.ppmi is True
    model.fit(src, tgt)
    assert [g.mult for g in model.optimizer.groups] == [1, 2]
    logits, labels = model.predict(tgt)
    assert logits.shape == (1200, 3) and torch.isfinite(logits).all()
