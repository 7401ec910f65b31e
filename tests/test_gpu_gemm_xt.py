"""First-layer GEMMs from the TILE-PACKED sparse input matrix (csrc/gemm_xt.cu: sub-tiles expanded into dense
swizzled bf16 operand tiles in shared memory, tcgen05 MMAs) against the dense split-bf16 kernel and fp64: `self.lin(x)` (pygda/nn/prop_gcn_conv.py:205), `x @ W` (cached_gcn_conv.py:130) and their weight
gradients; the dense rebuild (gda_unpack_tiles_f32); the estimator-level switch (ops.x_tiles)."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


def _sparse(rows, cols, density, seed, scale_cols=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(rows, cols, generator=g) + 0.5) * (torch.rand(rows, cols, generator=g) < density)
    if scale_cols:
        x = x * torch.logspace(-2, 2, cols)
    return x


def _tiles(x):
    from pygda_b200.data import PackedTiles
    return PackedTiles(x, pin=False)


@pytest.mark.parametrize("rows,cols,n,density", [
    (128, 64, 128, 0.07),          # one tile, one stage
    (1000, 700, 128, 0.07),        # row tail (1000 % 128), column tail (700 % 64), strip padding
    (4100, 6775, 128, 0.07),       # the config-2 layer-1 shape in K
    (19000, 1000, 128, 0.07),      # 149 row tiles: more CTAs than SMs, no K split in the dense kernel either
    (300, 1000, 5, 0.05),          # narrow output (classes)
    (513, 200, 256, 0.10),         # two output-column tiles
    (260, 4096, 64, 0.002),        # gaps >= 255: escape bytes, the slow decode path
    (256, 256, 128, 0.6),          # more than 256 entries per sub-tile: the slow decode path
])
@pytest.mark.parametrize("w_in_out", [False, True])
def test_forward_matches_fp64_and_the_dense_split_kernel(rows, cols, n, density, w_in_out):
    from pygda_b200 import ops
    x = _sparse(rows, cols, density, seed=rows + cols)
    g = torch.Generator().manual_seed(1)
    w = torch.randn((cols, n) if w_in_out else (n, cols), generator=g)
    t = _tiles(x.cuda()).view()
    out, _ = ops.xt_fwd(t, w.cuda(), w_in_out)
    ref64 = x.double() @ (w.double() if w_in_out else w.double().t())
    assert_close(out, ref64, 3e-5, "tile-packed forward vs fp64")
    if n >= 64 and cols >= 64:
        dense = ops.gemm_split(ops.Split(x.cuda()), ops.Split(w.cuda()), False, not w_in_out, rows, n, cols)
        assert_close(out, dense, 3e-5, "tile-packed vs dense split kernel")


@pytest.mark.parametrize("rows,cols,n,density", [
    (128, 128, 128, 0.07),
    (1000, 700, 128, 0.07),
    (20000, 6775, 128, 0.07),      # split over the rows (whole waves), feature tail
    (3000, 300, 5, 0.05),
    (700, 260, 200, 0.10),
    (1000, 4096, 64, 0.002),
    (512, 256, 128, 0.6),
])
@pytest.mark.parametrize("w_in_out", [False, True])
def test_weight_gradient_matches_fp64(rows, cols, n, density, w_in_out):
    from pygda_b200 import ops
    x = _sparse(rows, cols, density, seed=rows + cols + 1)
    g = torch.Generator().manual_seed(2)
    gy = torch.randn(rows, n, generator=g)
    t = _tiles(x.cuda()).view()
    dw, _ = ops.xt_dw(t, gy.cuda(), w_in_out)
    ref = x.double().t() @ gy.double()
    assert_close(dw, ref if w_in_out else ref.t(), 3e-5, "tile-packed weight gradient vs fp64")
    dw2, _ = ops.xt_dw(t, gy.cuda(), w_in_out)
    assert torch.equal(dw, dw2), "fixed-order reduction of the row splits: deterministic"


def test_column_scales_keep_fp32_accuracy():
    """Entries spanning four decades: the (hi, lo) split made in the expander keeps every product term."""
    from pygda_b200 import ops
    x = _sparse(600, 900, 0.07, seed=3, scale_cols=True)
    w = torch.randn(128, 900, generator=torch.Generator().manual_seed(4))
    out, _ = ops.xt_fwd(_tiles(x.cuda()).view(), w.cuda(), False)
    assert_close(out, x.double() @ w.double().t(), 3e-5, "scaled columns")


@pytest.mark.parametrize("rows,cols,density", [(1, 1, 1.0), (70, 130, 0.07), (1000, 700, 0.07), (260, 4096, 0.002),
                                               (64, 64, 1.0)])
def test_dense_rebuild_is_bit_exact(rows, cols, density):
    x = _sparse(rows, cols, density, seed=7)
    x[0, 0] = -0.0
    host = _tiles(x)
    assert torch.equal(host.decode().view(torch.int32), x.view(torch.int32))
    dev = _tiles(x.cuda())
    for k in ("_vals", "_codes", "_ptr", "_seg"):
        assert torch.equal(dev.tensors()[k].cpu(), host.tensors()[k]), f"device-built {k} differs from the host build"
    out = torch.full((rows, cols), 7.0, device="cuda")
    dev.view().unpack_into(out)
    assert torch.equal(out.cpu().view(torch.int32), x.view(torch.int32))


def test_staged_data_carries_the_packed_copy_and_the_first_layer_uses_it():
    """Data.pin_memory() keeps a sparse x tile-packed; .to(cuda) rebuilds the dense x bit for bit AND attaches the
    packed device copy; GraphConvFn multiplies from it: forward identical to the dense path, gradients to 1e-5."""
    from pygda_b200 import ops
    from pygda_b200.data import Data, PackedTiles
    from pygda_b200.graph import graph_for
    n, f, h = 3000, 2000, 128
    x = _sparse(n, f, 0.07, seed=11)
    g = torch.Generator().manual_seed(12)
    ei = torch.randint(0, n, (2, 9000), generator=g)
    d = Data(x=x, edge_index=ei, y=torch.zeros(n, dtype=torch.long)).pin_memory()
    assert isinstance(d.__dict__["_packed_x"], PackedTiles)
    on = d.to("cuda")
    assert torch.equal(on.x.cpu().view(torch.int32), x.view(torch.int32))
    assert ops.x_tiles(on.x, h) is on.x._gda_tiles
    edited = d.to("cuda")
    edited.x.mul_(2.0)                                       # in-place edit: the attached packed copy is stale now
    assert ops.x_tiles(edited.x, h) is None
    graph = graph_for(on.edge_index, n)
    w = torch.randn(h, f, generator=g).cuda().requires_grad_()
    b = torch.zeros(h, device="cuda", requires_grad=True)
    gy = torch.randn(n, h, generator=g).cuda()
    res = {}
    for enabled in (True, False):
        ops.XT_ENABLED = enabled
        try:
            w.grad = b.grad = None
            y = ops.graph_conv(on.x, w, b, graph, 2)
            y.backward(gy)
            res[enabled] = (y.detach().clone(), w.grad.clone(), b.grad.clone())
        finally:
            ops.XT_ENABLED = True
    assert_close(res[True][0], res[False][0], 3e-5, "forward: tile-packed vs dense (the dense kernel splits K here)")
    assert_close(res[True][1], res[False][1], 1e-5, "dW: tile-packed vs dense")
    assert_close(res[True][2], res[False][2], 1e-6, "bias gradient")


def test_resident_constant_features_are_packed_once():
    from pygda_b200 import ops
    x = _sparse(2000, 3000, 0.05, seed=21).cuda()
    assert ops.x_tiles(x, 128) is None                       # not marked constant: never packed behind the caller's back
    ops.mark_constant(x)
    t = ops.x_tiles(x, 128)
    assert t is not None and t is ops.x_tiles(x, 128) and t.vals.numel() == int((x != 0).sum())
    dense = ops.mark_constant(torch.randn(2000, 3000, device="cuda"))
    assert ops.x_tiles(dense, 128) is None                   # too dense: stays on the dense split path
    w = torch.randn(128, 3000, device="cuda", requires_grad=True)
    y = ops.linear(x, w)
    y.backward(torch.ones_like(y))
    assert_close(y, x.double() @ w.double().t(), 3e-5, "linear from the packed constant")
    assert_close(w.grad, torch.ones(2000, 128, device="cuda").double().t() @ x.double(), 3e-5, "its weight gradient")


def test_exponent_packed_staging_values_are_rebuilt_bit_for_bit():
    """The pinned staging form packs the exponents of the values (3.5 bytes per value); gda_unpack_values_f32 rebuilds
    every fp32 bit pattern on the device -- also inf / nan / denormals / negative zero and values outside the
    15-binade window (the escape list)."""
    from pygda_b200.data import PackedTiles
    x = _sparse(700, 900, 0.07, seed=31)
    x = x / x.sum(1, keepdim=True).clamp(min=1e-9)
    x[0, 0], x[0, 1], x[0, 2], x[0, 3], x[0, 4], x[1, 0] = float("inf"), float("nan"), -0.0, 1e-42, -3.5e20, -1e-4
    host = PackedTiles(x, pin=True)
    assert host.compressed and int(host.vmeta[1]) >= 5 and "_vals" not in host.tensors()
    assert host.nbytes < 0.93 * PackedTiles(x, pin=False).nbytes
    assert torch.equal(host.decompress_values().view(torch.int32), host.vals.view(torch.int32))
    dense = host.to_dense("cuda")
    torch.cuda.synchronize()
    assert torch.equal(dense._gda_tiles.vals.cpu().view(torch.int32), host.vals.view(torch.int32))
    assert torch.equal(dense.cpu().view(torch.int32), x.view(torch.int32))
