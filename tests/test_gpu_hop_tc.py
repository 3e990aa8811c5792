"""tcgen05 hop kernel (F = 64) against the exact-fp32 CUDA-core hop kernel on identical inputs.
The aggregation is bit-identical; the 64x64 filter product is 3xTF32 with fp32 accumulation, so the
results agree to ~1e-6 relative (tolerance below: 2e-6 of the row scale)."""
import pytest
import torch

from mswe_gnn_b200 import lib
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _csr(ei, n):
    return lib.csr_build(ei[0].contiguous(), ei[1].contiguous(), None, 0, n, 0, n)


BACKEND = ["tc"]          # "tc": 3xTF32 (swe_hop_tc.cu), "tc16": fp16 hi/lo splits (swe_hop_tc16.cu), "tc16s": + s-ring; set by the fixture below


@pytest.fixture(autouse=True, params=["tc", "tc16", "tc16s"])
def _backend(request):
    BACKEND[0] = request.param
    yield


def _run_both(o, s, rowptr, src, n, W, with_grad, addend, act, slope, dst_lo=0, n_dst=None):
    n_dst = n if n_dst is None else n_dst
    wt = torch.empty(64, 64, device=DEV)
    lib.pack_linear(W.contiguous(), 64, wt)
    ref = torch.zeros_like(o)
    out = torch.zeros_like(o)
    agg = torch.zeros_like(o)
    lib.propagate_hop_fwd(o, o, s, rowptr, src, dst_lo, n_dst, wt, with_grad, 0, addend, act, slope, ref, 64)
    if BACKEND[0] in ("tc16", "tc16s"):
        img = torch.empty(lib.hop_tc16_image_bytes(), dtype=torch.uint8, device=DEV)
        lib.hop_tc16_pack(W.contiguous(), float(W.abs().max()), img)
        (lib.propagate_hop_tc16s_fwd if BACKEND[0] == "tc16s" else lib.propagate_hop_tc16_fwd)(o, o, s, rowptr, src, dst_lo, n_dst, img, with_grad, 0, addend, act, slope, agg, out)
    else:
        img = torch.empty(lib.hop_tc_image_bytes(), dtype=torch.uint8, device=DEV)
        lib.hop_tc_pack(W.contiguous(), img)
        lib.propagate_hop_tc_fwd(o, o, s, rowptr, src, dst_lo, n_dst, img, with_grad, 0, addend, act, slope, agg, out)
    torch.cuda.synchronize()
    return ref, out, agg


@pytest.mark.parametrize("nx,ny,with_grad,act", [(5, 3, 1, 0), (40, 31, 1, 3), (64, 64, 0, 1), (131, 77, 1, 0)])
def test_hop_tc_matches_ffma_hop(nx, ny, with_grad, act):
    torch.manual_seed(nx)
    d = make_single_scale_mesh(nx, ny, seed=1)
    n, e = d.x.shape[0], d.edge_index.shape[1]
    rowptr, src, dst, eid = _csr(d.edge_index.to(DEV), n)
    o = torch.randn(n, 64, device=DEV)
    o[torch.rand(n, device=DEV) < 0.3] = 0
    s = torch.randn(e, 64, device=DEV)
    s = s / s.norm(dim=1, keepdim=True)
    W = torch.randn(64, 64, device=DEV) / 8
    addend = torch.randn(n, 64, device=DEV) if act else None
    slope = torch.tensor([0.25], device=DEV) if act == 1 else None
    ref, out, agg = _run_both(o, s, rowptr, src, n, W, with_grad, addend, act, slope)
    scale = float(ref.abs().max())
    assert float((out - ref).abs().max()) <= 2e-6 * scale, float((out - ref).abs().max()) / scale
    # the aggregation itself is exact fp32 in the reference's edge order
    row, col = d.edge_index.to(DEV)
    t = (o[col] - o[row]) * s[torch.argsort(eid.long())] if with_grad else s[torch.argsort(eid.long())] * o[row]
    expect = torch.zeros_like(o).index_add_(0, col, t)
    assert torch.allclose(agg, expect, rtol=1e-5, atol=1e-5)


def test_hop_tc_high_degree_and_row_range():
    """A star graph (one node with in-degree far above the staging cap) and a destination sub-range."""
    torch.manual_seed(0)
    n = 3000
    hub_edges = torch.stack([torch.arange(1, n), torch.zeros(n - 1, dtype=torch.long)])
    ring = torch.stack([torch.arange(n), (torch.arange(n) + 1) % n])
    ei = torch.cat([hub_edges, ring, ring.flip(0)], 1).to(DEV)
    e = ei.shape[1]
    rowptr, src, dst, eid = _csr(ei, n)
    o = torch.randn(n, 64, device=DEV)
    s = torch.randn(e, 64, device=DEV) / 30
    W = torch.randn(64, 64, device=DEV) / 8
    ref, out, _ = _run_both(o, s, rowptr, src, n, W, 1, None, 0, None)
    scale = float(ref.abs().max())
    assert float((out - ref).abs().max()) <= 4e-6 * scale
    # sub-range of destinations: rows outside stay untouched
    lo, cnt = 1000, 517
    rp2, src2, dst2, _ = lib.csr_build(ei[0].contiguous(), ei[1].contiguous(), None, 0, n, 0, n)
    sel = (dst2 >= lo) & (dst2 < lo + cnt)
    rp_sub = (rp2[lo:lo + cnt + 1] - rp2[lo]).contiguous()
    src_sub, s_sub = src2[sel].contiguous(), s[sel].contiguous()
    ref, out, _ = _run_both(o, s_sub, rp_sub, src_sub, n, W, 1, None, 0, None, dst_lo=lo, n_dst=cnt)
    assert float((out - ref).abs().max()) <= 4e-6 * scale
    assert float(out[:lo].abs().max()) == 0 and float(out[lo + cnt:].abs().max()) == 0


@pytest.mark.parametrize("mag", [1e-12, 1e-4, 1.0, 1e6])
def test_hop_tc_rows_of_any_magnitude(mag):
    """The fp16 kernel scales every agg row by its own power of two: node states of any magnitude (and rows that differ
    by many orders of magnitude inside one tile) keep the same relative accuracy."""
    torch.manual_seed(3)
    d = make_single_scale_mesh(40, 31, seed=2)
    n, e = d.x.shape[0], d.edge_index.shape[1]
    rowptr, src, dst, eid = _csr(d.edge_index.to(DEV), n)
    o = torch.randn(n, 64, device=DEV) * mag
    o[::7] *= 1e3
    o[3::11] *= 1e-3
    s = torch.randn(e, 64, device=DEV)
    s = s / s.norm(dim=1, keepdim=True)
    W = torch.randn(64, 64, device=DEV) / 8
    ref, out, _ = _run_both(o, s, rowptr, src, n, W, 1, None, 0, None)
    err = (out.double() - ref.double()).norm(dim=1) / ref.double().norm(dim=1).clamp_min(1e-300)
    assert bool(torch.isfinite(out).all())
    assert float(err.max()) < 5e-6, float(err.max())           # per ROW, not per tile


def test_hop_tc16s_bit_identical_to_tc16():
    """The s-ring edition changes how the gate rows reach the gather warps (cp.async.bulk into per-warp buffers, src ids
    and rowptr staged ahead), not the arithmetic or its order: equal bits on a mesh, on hubs that overflow the pass
    buffers (in-degree > 3 on a pass's four nodes) and the 32-id staging, and on isolated nodes."""
    if BACKEND[0] != "tc16s":
        pytest.skip("one run is enough")
    torch.manual_seed(11)
    d = make_single_scale_mesh(97, 53, seed=4)
    n = d.x.shape[0]
    hubs = torch.tensor([5, 130, 131, 132, 133, 1000, n - 1])
    hub_edges = torch.stack([torch.randint(0, n, (len(hubs) * 57,)), hubs.repeat_interleave(57)])
    keep = (d.edge_index[1] % 41) != 7                        # nodes 7, 48, 89, ... lose all their edges
    ei = torch.cat([d.edge_index[:, keep], hub_edges], 1).to(DEV)
    e = ei.shape[1]
    rowptr, src, dst, eid = _csr(ei, n)
    o = torch.randn(n, 64, device=DEV)
    s = torch.randn(e, 64, device=DEV) / 4
    W = torch.randn(64, 64, device=DEV) / 8
    img = torch.empty(lib.hop_tc16_image_bytes(), dtype=torch.uint8, device=DEV)
    lib.hop_tc16_pack(W.contiguous(), float(W.abs().max()), img)
    add = torch.randn(n, 64, device=DEV)
    slope = torch.tensor([0.25], device=DEV)
    for with_grad, addend, act in [(1, None, 0), (0, add, 1), (1, add, 3)]:
        a, b = torch.zeros_like(o), torch.zeros_like(o)
        agg_a, agg_b = torch.zeros_like(o), torch.zeros_like(o)
        lib.propagate_hop_tc16_fwd(o, o, s, rowptr, src, 0, n, img, with_grad, 0, addend, act, slope, agg_a, a)
        lib.propagate_hop_tc16s_fwd(o, o, s, rowptr, src, 0, n, img, with_grad, 0, addend, act, slope, agg_b, b)
        torch.cuda.synchronize()
        assert torch.equal(agg_a, agg_b)
        assert torch.equal(a, b)
    # repeated launches give the same bits (buffer hand-over races would show here)
    for _ in range(5):
        lib.propagate_hop_tc16s_fwd(o, o, s, rowptr, src, 0, n, img, 1, 0, add, 3, slope, agg_b, b)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_edge_set_block_bound_selects_the_hop():
    """plan.EdgeSet.max_block4 (edges of 4 consecutive destinations, as the s-ring hop cuts its passes) is what
    engine.SweGnnLauncher.run reads to keep graphs beyond the staging on the per-thread-load hop."""
    if BACKEND[0] != "tc16s":
        pytest.skip("one run is enough")
    from mswe_gnn_b200.plan import _build_edge_set
    d = make_single_scale_mesh(40, 31, seed=2)
    n = d.x.shape[0]
    ei = d.edge_index.to(DEV)
    es = _build_edge_set(ei[0], ei[1], None, 0, n, 0, n)
    deg = torch.bincount(ei[1], minlength=n)
    blocks = torch.nn.functional.pad(deg, (0, (-n) % 4)).view(-1, 4).sum(1)
    assert es.max_block4 == int(blocks.max()) <= 12                    # a dual mesh: in-degree <= 3
    hub = torch.stack([torch.arange(20, device=DEV), torch.full((20,), 7, device=DEV)])
    ei2 = torch.cat([ei, hub], 1)
    es2 = _build_edge_set(ei2[0].contiguous(), ei2[1].contiguous(), None, 0, n, 0, n)
    assert es2.max_block4 > 12
