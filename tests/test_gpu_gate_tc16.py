"""GPU: the tcgen05 edge-gate kernel with fp16 hi/lo splits (kind::f16, swe_gate_tc16.cu), stage by stage against fp64,
against the exact-fp32 CUDA-core gate kernel, and its range guard (tiles whose layer-0 inputs leave the fp16 window are
redone by the 3xTF32 kernel).  Tolerance: every layer within rel 1e-5 (north_star) of the fp64 result — the same bounds
as the 3xTF32 kernel's tests (test_gpu_gate_tc.py)."""
import pytest
import torch

import mswe_gnn_b200  # noqa: F401
from mswe_gnn_b200 import lib
from mswe_gnn_b200.engine import PackedGateTC, PackedMLP
from test_gpu_gate_tc import _setup

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run16(tc, xs, xd, xd_dst, a, src, dst, E, k1, s, dbg=None, guard=True):
    codes, slopes = tc.acts_and_slopes()
    ws = tc.flag_ws(E) if guard else None
    lib.edge_gate_tc16_fwd(xs, xd, xd_dst, a, src, dst, E, tc.image16(), tc.image() if guard else None, k1, codes, slopes,
                           True, s, dbg, ws)
    return ws


@pytest.mark.parametrize("n_edge_feat,drop_dst", [(64, False), (0, False), (0, True)])
def test_gate_tc16_stagewise_vs_fp64(n_edge_feat, drop_dst):
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(n_edge_feat)
    tc = PackedGateTC(mlp)
    s = torch.full((E, 64), float("nan"), device=DEV)
    dbg = torch.zeros(128 * 128 * 2 + 128 * 64, device=DEV)
    ws = _run16(tc, xs, xd, None if drop_dst else xd, a, src, dst, E, k1, s, dbg)
    torch.cuda.synchronize()
    assert int(ws[0]) == 0                                            # O(1) inputs: nothing leaves the window
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
    assert err1 < 3e-6, f"layer-0 accumulators (SS path, 64-byte swizzle) off by {err1:.3e}"
    h1 = torch.nn.functional.prelu(pre1 + lins[0].bias.double(), prl[0].weight.double()).detach()
    pre2 = (h1 @ lins[1].weight.double().T).detach()
    d2 = dbg[128 * 128:2 * 128 * 128].view(128, 128)[:T].double()
    err2 = float((d2 - pre2[:T]).abs().max() / pre2[:T].abs().max())
    assert err2 < 4e-6, f"layer-1 accumulators (TS path, fp16 A operand in TMEM) off by {err2:.3e}"
    h2 = torch.nn.functional.prelu(pre2 + lins[1].bias.double(), prl[1].weight.double()).detach()
    pre3 = (h2 @ lins[2].weight.double().T).detach()
    d3 = dbg[2 * 128 * 128:].view(128, 64)[:T].double()
    err3 = float((d3 - pre3[:T]).abs().max() / pre3[:T].abs().max())
    assert err3 < 6e-6, f"layer-2 accumulators (TS, N=64) off by {err3:.3e}"
    u = torch.nn.functional.prelu(pre3 + lins[2].bias.double(), prl[2].weight.double()).detach()
    ref = u / u.norm(dim=1, keepdim=True)
    err = float((s.double() - ref).abs().max())
    print(f"tc16 stage errors {err1:.2e} {err2:.2e} {err3:.2e} final {err:.2e}")
    assert err < 1e-5, f"normalised gate off by {err:.3e} (all tiles)"


@pytest.mark.parametrize("n_edge_feat,drop_dst,nx,ny", [(64, False, 60, 40), (0, False, 33, 21), (0, True, 33, 21), (64, False, 5, 3),
                                                      (0, True, 400, 300), (64, False, 400, 300)])   # the last two: ~19 tiles per CTA
def test_gate_tc16_matches_exact_fp32_kernel(n_edge_feat, drop_dst, nx, ny):
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(n_edge_feat, seed=3, nx=nx, ny=ny)
    tc = PackedGateTC(mlp)
    xd_dst = None if drop_dst else xd
    pad = 4096
    buf = torch.full((E * 64 + 2 * pad,), 7.0, device=DEV)
    s16 = buf[pad:pad + E * 64].view(E, 64)
    ws = _run16(tc, xs, xd, xd_dst, a, src, dst, E, k1, s16)
    s_ff = torch.empty(E, 64, device=DEV)
    pk = PackedMLP(mlp, [(64, 64)] * (5 if n_edge_feat else 4), {})
    lib.edge_gate_fwd(xs, xd, xd_dst, a, src, dst, E, pk.struct(), True, s_ff, 64)
    torch.cuda.synchronize()
    assert int(ws[0]) == 0
    assert bool(torch.isfinite(s16).all())
    assert float((s16 - s_ff).abs().max()) < 1e-5, float((s16 - s_ff).abs().max())
    assert bool((buf[:pad] == 7).all()) and bool((buf[pad + E * 64:] == 7).all())          # nothing written outside
    s2 = torch.empty(E, 64, device=DEV)
    _run16(tc, xs, xd, xd_dst, a, src, dst, E, k1, s2)
    assert torch.equal(s16, s2)                                                           # deterministic


@pytest.mark.parametrize("what", ["huge", "tiny", "inf_free_mixed"])
def test_gate_tc16_range_guard_falls_back_to_tf32(what):
    """Layer-0 inputs outside the fp16 window: rows above 2^15 (would overflow fp16) or entirely
    below 2^-5 (would lose relative precision) are listed and redone by the 3xTF32 kernel; the result stays within the
    same 1e-5 of the exact-fp32 kernel and the list is non-empty."""
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(64, seed=9, nx=60, ny=40)
    if what == "huge":
        xs = xs.clone(); xs[: n // 3] *= 3.0e5
    elif what == "tiny":
        xs, xd, a = xs * 1e-5, xd * 1e-5, a * 1e-5
    else:
        xd = xd.clone(); xd[n // 2:] *= 5.0e4
    tc = PackedGateTC(mlp)
    s16 = torch.full((E, 64), float("nan"), device=DEV)
    ws = _run16(tc, xs, xd, xd, a, src, dst, E, k1, s16)
    s_ff = torch.empty(E, 64, device=DEV)
    pk = PackedMLP(mlp, [(64, 64)] * 5, {})
    lib.edge_gate_fwd(xs, xd, xd, a, src, dst, E, pk.struct(), True, s_ff, 64)
    torch.cuda.synchronize()
    n_tiles = (E + 127) // 128
    assert 0 < int(ws[0]) <= n_tiles, int(ws[0])
    listed = ws[1:1 + int(ws[0])]
    assert bool(((listed >= 0) & (listed < n_tiles)).all()) and listed.unique().numel() == listed.numel()
    assert bool(torch.isfinite(s16).all())
    assert float((s16 - s_ff).abs().max()) < 1e-5, float((s16 - s_ff).abs().max())


def test_gate_tc16_hidden_rows_of_any_magnitude():
    """The hidden activations are scaled per row, so weights that make them huge or tiny change nothing."""
    for scale in (1e-6, 1.0, 1e6):
        n, E, src, dst, xs, xd, a, mlp, k1 = _setup(64, seed=13, nx=33, ny=21)
        with torch.no_grad():
            lins = [m for m in mlp if isinstance(m, torch.nn.Linear)]
            lins[0].weight.mul_(scale); lins[0].bias.mul_(scale)
        tc = PackedGateTC(mlp)
        s16 = torch.full((E, 64), float("nan"), device=DEV)
        ws = _run16(tc, xs, xd, xd, a, src, dst, E, k1, s16)
        s_ff = torch.empty(E, 64, device=DEV)
        pk = PackedMLP(mlp, [(64, 64)] * 5, {})
        lib.edge_gate_fwd(xs, xd, xd, a, src, dst, E, pk.struct(), True, s_ff, 64)
        torch.cuda.synchronize()
        assert int(ws[0]) == 0
        assert float((s16 - s_ff).abs().max()) < 1e-5, (scale, float((s16 - s_ff).abs().max()))


def test_gate_tc16_generic_activation_and_zero_rows():
    """tanh edge MLP (the out-of-line activation path) and an all-zero output row (0/0 -> NaN -> 0, gnn.py:425-426)."""
    from mswe_gnn_b200.models.models import make_mlp
    n, E, src, dst, xs, xd, a, _, k1 = _setup(64, seed=2, nx=33, ny=21)
    torch.manual_seed(4)
    mlp = make_mlp(k1, 64, hidden_size=128, n_layers=3, bias=True, activation="tanh").to(DEV)
    tc = PackedGateTC(mlp)
    s16 = torch.full((E, 64), float("nan"), device=DEV)
    _run16(tc, xs, xd, xd, a, src, dst, E, k1, s16)
    s_ff = torch.empty(E, 64, device=DEV)
    pk = PackedMLP(mlp, [(64, 64)] * 5, {})
    lib.edge_gate_fwd(xs, xd, xd, a, src, dst, E, pk.struct(), True, s_ff, 64)
    torch.cuda.synchronize()
    assert float((s16 - s_ff).abs().max()) < 1e-5, float((s16 - s_ff).abs().max())
    n, E, src, dst, xs, xd, a, mlp, k1 = _setup(0, seed=5)
    with torch.no_grad():
        lins = [m for m in mlp if isinstance(m, torch.nn.Linear)]
        lins[2].weight.zero_(); lins[2].bias.zero_()
    tc = PackedGateTC(mlp)
    s = torch.full((E, 64), float("nan"), device=DEV)
    _run16(tc, xs, xd, xd, None, src, dst, E, k1, s)
    assert float(s.abs().max()) == 0.0
    # zero edges: nothing is touched
    s = torch.full((4, 64), 2.0, device=DEV)
    lib.edge_gate_tc16_fwd(xs, xd, xd, None, src, dst, 0, tc.image16(), tc.image(), k1, *tc.acts_and_slopes(), True, s, None,
                           tc.flag_ws(1))
    assert bool((s == 2.0).all())
