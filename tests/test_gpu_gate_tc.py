"""GPU: the tcgen05 (3xTF32) edge-gate kernel, stage by stage against fp64 and against the exact-fp32
CUDA-core gate kernel.  Tolerance: every layer within rel 1e-5 (north_star) of the fp64 result."""
import os

import pytest
import torch

import mswe_gnn_b200  # noqa: F401
from mswe_gnn_b200 import lib
from mswe_gnn_b200.engine import PackedGateTC, PackedMLP
from mswe_gnn_b200.models.models import make_mlp
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(n_edge_feat, seed=0, nx=12, ny=9):
    torch.manual_seed(seed)
    d = make_single_scale_mesh(nx, ny, seed=seed)
    n = d.x.shape[0]
    row, col = d.edge_index.to(DEV)
    rowptr, src, dst, eid = lib.csr_build(row, col, None, 0, n, 0, n)
    E = int(src.numel())
    xs = torch.randn(n, 64, device=DEV)
    xd = torch.randn(n, 64, device=DEV) * (torch.rand(n, 1, device=DEV) < 0.6)
    a = torch.randn(E, 64, device=DEV) if n_edge_feat else None
    k1 = 256 + n_edge_feat
    mlp = make_mlp(k1, 64, hidden_size=128, n_layers=3, bias=True, activation="prelu").to(DEV)
    with torch.no_grad():
        for m in mlp:
            if isinstance(m, torch.nn.PReLU):
                m.weight.fill_(0.1 + 0.2 * torch.rand(1).item())
    return n, E, src, dst, xs, xd, a, mlp, k1


@pytest.mark.parametrize("n_edge_feat,drop_dst", [(64, False), (0, False), (0, True)])
def test_gate_tc_stagewise_vs_fp64(n_edge_feat, drop_dst):
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(n_edge_feat)
    tc = PackedGateTC(mlp)
    assert PackedGateTC.eligible(mlp, 64)
    codes, slopes = tc.acts_and_slopes()
    s = torch.full((E, 64), float("nan"), device=DEV)
    dbg = torch.zeros(128 * 128 * 2 + 128 * 64, device=DEV)
    lib.edge_gate_tc_fwd(xs, xd, None if drop_dst else xd, a, src, dst, E, tc.image(), k1, codes, slopes, True, s, dbg)
    torch.cuda.synchronize()
    # fp64 reference of the first tile, stage by stage
    sl, dl = src.long(), dst.long()
    xdd = torch.zeros_like(xd) if drop_dst else xd
    parts = [xs[sl], xs[dl], xd[sl], xdd[dl]] + ([a] if a is not None else [])
    z = torch.cat(parts, 1).double()
    lins = [m for m in mlp if isinstance(m, torch.nn.Linear)]
    prl = [m for m in mlp if isinstance(m, torch.nn.PReLU)]
    T = min(128, E)
    pre1 = (z @ lins[0].weight.double().T).detach()
    d1 = dbg[:128 * 128].view(128, 128)[:T].double()
    err1 = float((d1 - pre1[:T]).abs().max() / pre1[:T].abs().max())
    print('stage errors:', err1); assert err1 < 3e-6, f"layer-0 accumulators (SS path) off by {err1:.3e}"
    h1 = torch.nn.functional.prelu(pre1 + lins[0].bias.double(), prl[0].weight.double()).detach()
    pre2 = (h1 @ lins[1].weight.double().T).detach()
    d2 = dbg[128 * 128:2 * 128 * 128].view(128, 128)[:T].double()
    err2 = float((d2 - pre2[:T]).abs().max() / pre2[:T].abs().max())
    print('stage2', err2); assert err2 < 4e-6, f"layer-1 accumulators (TS path) off by {err2:.3e}"
    h2 = torch.nn.functional.prelu(pre2 + lins[1].bias.double(), prl[1].weight.double()).detach()
    pre3 = (h2 @ lins[2].weight.double().T).detach()
    d3 = dbg[2 * 128 * 128:].view(128, 64)[:T].double()
    err3 = float((d3 - pre3[:T]).abs().max() / pre3[:T].abs().max())
    print('stage3', err3); assert err3 < 6e-6, f"layer-2 accumulators (TS, N=64) off by {err3:.3e}"
    u = torch.nn.functional.prelu(pre3 + lins[2].bias.double(), prl[2].weight.double()).detach()
    ref = u / u.norm(dim=1, keepdim=True)
    err = float((s.double() - ref).abs().max()); print('final', err)
    assert err < 1e-5, f"normalised gate off by {err:.3e} (all tiles)"


def test_gate_tc_matches_exact_fp32_kernel_on_many_tiles():
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(64, seed=3, nx=60, ny=40)      # 14k edges -> 111 tiles
    tc = PackedGateTC(mlp)
    codes, slopes = tc.acts_and_slopes()
    s_tc = torch.empty(E, 64, device=DEV)
    s_ff = torch.empty(E, 64, device=DEV)
    lib.edge_gate_tc_fwd(xs, xd, xd, a, src, dst, E, tc.image(), k1, codes, slopes, True, s_tc, None)
    pk = PackedMLP(mlp, [(64, 64)] * 5, {})
    lib.edge_gate_fwd(xs, xd, xd, a, src, dst, E, pk.struct(), True, s_ff, 64)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(s_tc).all())
    assert float((s_tc - s_ff).abs().max()) < 1e-5
    # deterministic
    s2 = torch.empty_like(s_tc)
    lib.edge_gate_tc_fwd(xs, xd, xd, a, src, dst, E, tc.image(), k1, codes, slopes, True, s2, None)
    assert torch.equal(s_tc, s2)


def test_gate_tc_zero_rows_give_zero_gate_not_nan():
    """An all-zero MLP output row normalises to 0/0 -> NaN -> 0 (gnn.py:425-426)."""
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(0, seed=5)
    with torch.no_grad():
        lins = [m for m in mlp if isinstance(m, torch.nn.Linear)]
        lins[2].weight.zero_(); lins[2].bias.zero_()
    tc = PackedGateTC(mlp)
    codes, slopes = tc.acts_and_slopes()
    s = torch.full((E, 64), float("nan"), device=DEV)
    lib.edge_gate_tc_fwd(xs, xd, xd, None, src, dst, E, tc.image(), k1, codes, slopes, True, s, None)
    assert float(s.abs().max()) == 0.0


@pytest.mark.parametrize("n_edge_feat,drop_dst,nx,ny", [(64, False, 60, 40), (0, False, 33, 21), (0, True, 33, 21), (64, False, 5, 3),
                                                      (0, True, 400, 300), (64, False, 400, 300)])   # the last two: ~19 tiles per CTA
def test_gate_tc_decomposed_layer0_matches_exact_fp32_kernel(n_edge_feat, drop_dst, nx, ny):
    """Per-node partial tables (swe_gate_partials_tc) + decomposed gate against the exact-fp32 CUDA-core gate and the
    full tcgen05 gate: the first layer is only re-associated (P_src[r] + P_dst[c] + E·a)."""
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(n_edge_feat, seed=7, nx=nx, ny=ny)
    tc = PackedGateTC(mlp)
    codes, slopes = tc.acts_and_slopes()
    xd_dst = None if drop_dst else xd
    nseg = 3 + (0 if drop_dst else 1) + (1 if n_edge_feat else 0)
    # fp64 table check:  P_src = A xs + C xd,  P_dst = B xs + D xd
    W1 = [m for m in mlp if isinstance(m, torch.nn.Linear)][0].weight.detach().double()
    p_src = torch.full((n, 128), float("nan"), device=DEV)
    p_dst = torch.full((n, 128), float("nan"), device=DEV)
    lib.gate_partials_tc(xs, xd, 0, n, tc.image(), k1, 0, p_src)
    lib.gate_partials_tc(xs, xd_dst, 0, n, tc.image(), k1, 1, p_dst)
    ref_src = xs.double() @ W1[:, 0:64].T + xd.double() @ W1[:, 128:192].T
    ref_dst = xs.double() @ W1[:, 64:128].T + (0 if drop_dst else xd.double() @ W1[:, 192:256].T)
    scale = float(ref_src.abs().max())
    assert float((p_src.double() - ref_src).abs().max()) < 2e-6 * scale
    assert float((p_dst.double() - ref_dst).abs().max()) < 2e-6 * scale
    s_dec = torch.empty(E, 64, device=DEV)
    lib.edge_gate_tc_dec_fwd(p_src, p_dst, a, src, dst, E, tc.image(), k1, codes, slopes, True, s_dec)
    s_ff = torch.empty(E, 64, device=DEV)
    pk = PackedMLP(mlp, [(64, 64)] * (5 if n_edge_feat else 4), {})
    lib.edge_gate_fwd(xs, xd, xd_dst, a, src, dst, E, pk.struct(), True, s_ff, 64)
    s_full = torch.empty(E, 64, device=DEV)
    lib.edge_gate_tc_fwd(xs, xd, xd_dst, a, src, dst, E, tc.image(), k1, codes, slopes, True, s_full, None)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(s_dec).all())
    assert float((s_dec - s_ff).abs().max()) < 1e-5, float((s_dec - s_ff).abs().max())
    assert float((s_dec - s_full).abs().max()) < 1e-5
    # a sub-range of table rows leaves the other rows untouched
    p2 = torch.zeros(n, 128, device=DEV)
    lo, cnt = n // 3, n // 2
    lib.gate_partials_tc(xs, xd, lo, cnt, tc.image(), k1, 0, p2)
    assert torch.equal(p2[lo:lo + cnt], p_src[lo:lo + cnt]) and float(p2[:lo].abs().max()) == 0 and float(p2[lo + cnt:].abs().max()) == 0


@pytest.mark.parametrize("hoist_xs", [False, True])
@pytest.mark.parametrize("n_edge_feat,drop_dst,nx,ny", [(64, False, 60, 40), (0, False, 33, 21), (0, True, 33, 21), (64, False, 5, 3),
                                                      (0, True, 400, 300), (64, False, 400, 300)])   # the last two: ~19 tiles per CTA
def test_gate_tc_static_share_hoisted(n_edge_feat, drop_dst, nx, ny, hoist_xs):
    """Per-edge table of the static share of layer 0 (a_e; with hoist_xs also x_s[r], x_s[c] — constant over a rollout
    of a with_WL=False model) + the per-step gate on the remaining blocks, against fp64, the exact-fp32 CUDA-core
    gate and the full tcgen05 gate."""
    if not hoist_xs and not n_edge_feat:
        pytest.skip("nothing to hoist")
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(n_edge_feat, seed=11, nx=nx, ny=ny)
    tc = PackedGateTC(mlp)
    codes, slopes = tc.acts_and_slopes()
    xd_dst = None if drop_dst else xd
    W1 = [m for m in mlp if isinstance(m, torch.nn.Linear)][0].weight.detach().double()
    n_tiles = (E + 127) // 128
    tab = torch.full((n_tiles * 128, 128), float("nan"), device=DEV)
    lib.gate_static_partials_tc(xs if hoist_xs else None, a, src, dst, E, tc.image(), k1, tab)
    ref = torch.zeros(E, 128, dtype=torch.float64, device=DEV)
    if hoist_xs:
        ref = ref + xs.double()[src.long()] @ W1[:, 0:64].T + xs.double()[dst.long()] @ W1[:, 64:128].T
    if n_edge_feat:
        ref = ref + a.double() @ W1[:, 256:320].T
    # internal order [tile][column half][16-byte chunk][row][4] -> [edge][128]
    rowmajor = tab.view(n_tiles, 2, 16, 128, 4).permute(0, 3, 1, 2, 4).reshape(n_tiles * 128, 128)[:E]
    assert float((rowmajor.double() - ref).abs().max()) < 2e-6 * float(ref.abs().max())
    s_st = torch.full((E, 64), float("nan"), device=DEV)
    xs_step = None if hoist_xs else xs
    lib.edge_gate_tc_stat_fwd(tab, xs_step, xd, xd_dst, src, dst, E, tc.image(), k1, codes, slopes, True, s_st)
    s_ff = torch.empty(E, 64, device=DEV)
    pk = PackedMLP(mlp, [(64, 64)] * (5 if n_edge_feat else 4), {})
    lib.edge_gate_fwd(xs, xd, xd_dst, a, src, dst, E, pk.struct(), True, s_ff, 64)
    s_full = torch.empty(E, 64, device=DEV)
    lib.edge_gate_tc_fwd(xs, xd, xd_dst, a, src, dst, E, tc.image(), k1, codes, slopes, True, s_full, None)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(s_st).all())
    assert float((s_st - s_ff).abs().max()) < 1e-5, float((s_st - s_ff).abs().max())
    assert float((s_st - s_full).abs().max()) < 1e-5
    # deterministic
    s2 = torch.empty(E, 64, device=DEV)
    lib.edge_gate_tc_stat_fwd(tab, xs_step, xd, xd_dst, src, dst, E, tc.image(), k1, codes, slopes, True, s2)
    torch.cuda.synchronize()
    assert torch.equal(s2, s_st)


@pytest.mark.parametrize("with_WL", [True, False])
def test_rollout_with_hoisted_static_share_matches_full_gate(monkeypatch, with_WL):
    """rollout_test (CUDA-graph loop) with the static share hoisted (default) against MSWE_GATE_L0=full, and a second
    rollout from other static columns through RolloutRunner.reset(x): the tables must be rebuilt.  with_WL=True
    (config.yaml): x_s is encoded from the current water level, only a_e is hoisted; the boundary inflow below makes
    the water level change by O(1) over the steps, so a table wrongly built from x_s would show."""
    from helpers import REF_CONFIG_MODELS, rel_l2
    from mswe_gnn_b200.models.gnn import MSGNN
    from mswe_gnn_b200.training.train import RolloutRunner
    from mswe_gnn_b200.utils.synthetic import make_tri_mesh
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **REF_CONFIG_MODELS)
    ctor["with_WL"] = with_WL
    m = MSGNN(**ctor).to(DEV).eval()
    with torch.no_grad():                                            # a decoder that moves the water depth a lot
        for p_ in m.node_decoder.parameters():
            p_.mul_(4.0)
    d = make_tri_mesh(40, 32, 3, rollout_steps=6, seed=3).to(DEV)
    x2 = d.x.clone()
    x2[:, :2] = torch.randn_like(x2[:, :2]) * 3.0                   # other static features, same dynamic state
    out = {}
    monkeypatch.setenv("MSWE_GATE", "tc")                            # the hoisting belongs to the 3xTF32 kernel
    with torch.no_grad():
        for mode in ("static", "full"):
            monkeypatch.setenv("MSWE_GATE_L0", mode)
            r = RolloutRunner(m, d, 6)
            a = r.run().clone()
            r.reset(x2)
            b = r.run().clone()
            out[mode] = (a, b)
    for i in range(2):
        e = rel_l2(out["static"][i], out["full"][i])
        assert e < 5e-5, (i, e)                                     # 6 autoregressive steps: bounded drift
    e = rel_l2(out["full"][0], out["full"][1])
    assert e > 1e-4, e                                               # the two rollouts really differ
    depth = out["full"][0][:, :, 0]
    assert float((depth[-1] - depth[0]).abs().max()) > 1e-2          # and the water level really moves


@pytest.mark.parametrize("nx,ny", [(5, 3), (60, 40), (129, 77)])
def test_gate_training_forward_and_static_table_stay_inside_their_buffers(nx, ny):
    """Guard bands around every output of the training forward (s, three pre-activations, work lists) and of the static
    table producer (padded to whole tiles by contract): nothing is written outside (no compute-sanitizer on this pool)."""
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(64, seed=5, nx=nx, ny=ny)
    tc = PackedGateTC(mlp)
    codes, slopes = tc.acts_and_slopes()
    pad = 4096

    def guarded(n_elems, dtype=torch.float32):
        buf = torch.full((n_elems + 2 * pad,), 7, device=DEV, dtype=dtype)
        return buf, buf[pad:pad + n_elems]

    bufs = {}
    for name, ne in (("s", E * 64), ("p1", E * 128), ("p2", E * 128), ("p3", E * 64)):
        bufs[name] = guarded(ne)
    cap = 1 << 12
    lists = guarded(3 * cap, torch.int64)
    count = torch.zeros(4, dtype=torch.int32, device=DEV)
    lib.edge_gate_tc_train_fwd(xs, xd, xd, a, src, dst, E, tc.image(), k1, codes, slopes, True, bufs["p1"][1].view(E, 128),
                               bufs["p2"][1].view(E, 128), bufs["p3"][1].view(E, 64), bufs["s"][1].view(E, 64), lists[1], count,
                               cap, 1e-2)                       # a wide threshold: the lists overflow their capacity
    n_tiles = (E + 127) // 128
    tab = guarded(n_tiles * 128 * 128)
    lib.gate_static_partials_tc(xs, a, src, dst, E, tc.image(), k1, tab[1])
    torch.cuda.synchronize()
    for buf, view in list(bufs.values()) + [lists, tab]:
        assert bool((buf[:pad] == 7).all()) and bool((buf[pad + view.numel():] == 7).all())
    assert int(count[:3].max()) > 0
    # the listed entries are in range even when the lists overflow
    for layer, width in ((0, 128), (1, 128), (2, 64)):
        k = min(int(count[layer]), cap)
        ent = lists[1][layer * cap: layer * cap + k]
        assert bool(((ent >> 8) < E).all()) and bool(((ent & 255) < width).all())


def test_gate_static_and_training_entry_points_accept_zero_edges():
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(64, seed=1, nx=5, ny=3)
    tc = PackedGateTC(mlp)
    codes, slopes = tc.acts_and_slopes()
    s = torch.full((4, 64), 2.0, device=DEV)
    tab = torch.full((128, 128), 2.0, device=DEV)
    p1, p2, p3 = (torch.full((4, w), 2.0, device=DEV) for w in (128, 128, 64))
    lib.gate_static_partials_tc(xs, a, src, dst, 0, tc.image(), k1, tab)
    lib.edge_gate_tc_stat_fwd(tab, None, xd, xd, src, dst, 0, tc.image(), k1, codes, slopes, True, s)
    lib.edge_gate_tc_train_fwd(xs, xd, xd, a, src, dst, 0, tc.image(), k1, codes, slopes, True, p1, p2, p3, s)
    torch.cuda.synchronize()
    for t in (s, tab, p1, p2, p3):
        assert bool((t == 2.0).all())
