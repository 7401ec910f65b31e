"""gda_graph_create vs the oracle's gcn_norm: indices bit-exact, weights ~1 ulp."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import pyg_ops as P

pytestmark = pytest.mark.gpu


def _graph(ei, n, w=None, flags=None):
    from pygda_b200.graph import Graph, SELF_LOOPS, NORM_SYM_COL
    flags = SELF_LOOPS | NORM_SYM_COL if flags is None else flags
    return Graph(ei.cuda(), n, None if w is None else w.cuda(), flags)


def test_golden_gcn_norm_cases_bit_exact_indices():
    from pygda_b200.graph import SELF_LOOPS, NORM_SYM_COL, IMPROVED
    g = load_golden("gcn_norm")
    flags = {"plain": SELF_LOOPS | NORM_SYM_COL, "improved": SELF_LOOPS | NORM_SYM_COL | IMPROVED,
             "weighted": SELF_LOOPS | NORM_SYM_COL, "noloops": NORM_SYM_COL}
    for name, case in g["cases"].items():
        w = g["edge_weight"] if name == "weighted" else None
        gr = _graph(g["edge_index"], g["num_nodes"], w, flags[name])
        ei, ew = gr.coo()
        assert torch.equal(ei.cpu(), case["edge_index_out"]), name
        assert_close(ew, case["weight_out"], 1e-6, name)


def test_golden_cached_norm_by_row():
    from pygda_b200.graph import SELF_LOOPS, NORM_SYM_ROW
    g = load_golden("cached_norm")
    gr = _graph(g["edge_index"], g["num_nodes"], None, SELF_LOOPS | NORM_SYM_ROW)
    ei, ew = gr.coo()
    assert torch.equal(ei.cpu(), g["edge_index_out"])
    assert_close(ew, g["weight_out"], 1e-6, "by-row weights")


@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (5, 0, 1), (300, 4000, 2), (5000, 60000, 3)])
def test_csr_matches_oracle_coo(n, e, seed):
    gen = torch.Generator().manual_seed(seed)
    ei = torch.randint(n, (2, e), generator=gen)
    gr = _graph(ei, n)
    ref_ei, ref_w = P.gcn_norm_by_col(ei, None, n)
    out_ei, out_w = gr.coo()
    assert torch.equal(out_ei.cpu(), ref_ei)
    assert_close(out_w, ref_w, 1e-6, "weights") if ref_w.numel() else None
    for transpose in (False, True):
        rp, ci, v = [t.cpu() for t in gr.csr(transpose)]
        key = ref_ei[0] if transpose else ref_ei[1]
        other = ref_ei[1] if transpose else ref_ei[0]
        order = torch.argsort(key, stable=True)
        counts = torch.bincount(key, minlength=n)
        assert torch.equal(rp.long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
        assert torch.equal(ci.long(), other[order])          # stable: COO order within a row
        assert_close(v, ref_w[order], 1e-6, "csr vals") if v.numel() else None


def test_out_of_range_index_is_an_error():
    from pygda_b200 import GdaError
    with pytest.raises(GdaError, match="outside"):
        _graph(torch.tensor([[0, 7], [1, 2]]), 4)


def test_long_rows_are_split():
    n = 2000
    hub = torch.zeros(1500, dtype=torch.long)
    leaves = torch.arange(1, 1501)
    ei = torch.stack([torch.cat([hub, leaves]), torch.cat([leaves, hub])])
    gr = _graph(ei, n)
    assert gr.num_long_rows == 1 and gr.num_long_rows_t == 1
