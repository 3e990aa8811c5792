"""tcgen05 forms of the two backward GEMMs of a Linear (dx = delta·W, dW = deltaᵀ·X) against fp64 torch matmuls and
the exact-fp32 CUDA-core kernels.  3xTF32 with fp32 accumulation: ~1e-6 relative (tolerance 1e-5 of the matrix norm)."""
import pytest
import torch

import mswe_gnn_b200  # noqa: F401
from mswe_gnn_b200 import lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _act_code(name):
    return lib.ACT_CODES[None if name == "none" else name]


@pytest.mark.parametrize("n_rows", [1, 97, 128, 4000, 150_001])
@pytest.mark.parametrize("n,ko,split", [(128, 128, 128), (128, 64, 64), (64, 128, 128), (128, 128, 64)])
def test_dx_tc_matches_fp64(n_rows, n, ko, split):
    g = torch.Generator(device="cpu").manual_seed(n_rows * 7 + n + ko)
    delta = torch.randn(n_rows, n, generator=g).to(DEV)
    w_ld = 320
    w = (torch.randn(n, w_ld, generator=g) * 0.1).to(DEV)
    k_off, k_valid = 64, ko
    ref = (delta.double() @ w.double()[:, k_off:k_off + ko]).float()
    if split == ko:
        dx = torch.full((n_rows, ko), float("nan"), device=DEV)
        lib.mlp_layer_bwd_dx_tc(delta, n_rows, n, w, w_ld, k_off, k_valid, ko, dx, False)
        torch.cuda.synchronize()
        assert _rel(dx, ref) < 1e-5
        # accumulate: a second pass doubles it
        lib.mlp_layer_bwd_dx_tc(delta, n_rows, n, w, w_ld, k_off, k_valid, ko, dx, True)
        torch.cuda.synchronize()
        assert _rel(dx, 2 * ref) < 1e-5
    else:
        a = torch.full((n_rows, 64), float("nan"), device=DEV)
        b = torch.ones(n_rows, 64, device=DEV)
        lib.mlp_layer_bwd_dx_tc(delta, n_rows, n, w, w_ld, k_off, k_valid, ko, a, False, b, True, 64)
        torch.cuda.synchronize()
        assert _rel(a, ref[:, :64]) < 1e-5
        assert _rel(b, ref[:, 64:] + 1) < 1e-5


def test_dx_tc_zero_pads_invalid_columns():
    n_rows, n, ko = 300, 128, 64
    delta = torch.randn(n_rows, n, device=DEV)
    w = torch.randn(n, 100, device=DEV)
    dx = torch.empty(n_rows, ko, device=DEV)
    lib.mlp_layer_bwd_dx_tc(delta, n_rows, n, w, 100, 60, 40, ko, dx, False)          # only 40 valid columns
    ref = torch.zeros(n_rows, ko, device=DEV, dtype=torch.float64)
    ref[:, :40] = delta.double() @ w.double()[:, 60:100]
    torch.cuda.synchronize()
    assert _rel(dx, ref) < 1e-5
    assert (dx[:, 40:] == 0).all()


def _reduce(part, grid, n, widths):
    """Sum the per-CTA partials [grid][Σ n*w] -> list of [n, w] blocks (fp64 on the host side of the test)."""
    tot = n * sum(widths)
    p = part[: grid * tot].view(grid, tot).double().sum(0)
    out, c = [], 0
    for w in widths:
        out.append(p[c * n: c * n + n * w].view(n, w))
        c += w
    return out


@pytest.mark.parametrize("n_rows", [1, 31, 32, 4097, 200_003])
def test_dw_tc_gathered_segments(n_rows):
    """n = 128, provider = gathered node rows (two index maps) + a dense edge block: layer 0 of the edge MLP."""
    g = torch.Generator(device="cpu").manual_seed(n_rows)
    n_nodes = max(n_rows // 3, 4)
    delta = torch.randn(n_rows, 128, generator=g).to(DEV)
    xs = torch.randn(n_nodes, 64, generator=g).to(DEV)
    xd = torch.randn(n_nodes, 64, generator=g).to(DEV)
    a = torch.randn(n_rows, 64, generator=g).to(DEV)
    src = torch.randint(0, n_nodes, (n_rows,), generator=g).to(torch.int32).to(DEV)
    dst = torch.randint(0, n_nodes, (n_rows,), generator=g).to(torch.int32).to(DEV)
    rows = lib.make_rows([(xs, src, 64, 64, 0, None), (xs, dst, 64, 64, 0, None), (xd, src, 64, 64, 0, None),
                          (xd, dst, 64, 64, 0, None)])
    grid = lib.mlp_layer_bwd_dw_tc_grid(n_rows)
    part = torch.full((grid * 128 * 256,), float("nan"), device=DEV)
    assert lib.mlp_layer_bwd_dw_tc(delta, n_rows, 128, rows, part) == grid
    torch.cuda.synchronize()
    blocks = _reduce(part, grid, 128, [64, 64, 64, 64])
    X = [xs[src.long()], xs[dst.long()], xd[src.long()], xd[dst.long()]]
    for b, x in zip(blocks, X):
        assert _rel(b, delta.double().t() @ x.double()) < 1e-5
    # the remaining 64-wide block on its own
    rows = lib.make_rows([(a, None, 64, 64, 0, None)])
    part = torch.full((grid * 128 * 64,), float("nan"), device=DEV)
    lib.mlp_layer_bwd_dw_tc(delta, n_rows, 128, rows, part)
    torch.cuda.synchronize()
    (b,) = _reduce(part, grid, 128, [64])
    assert _rel(b, delta.double().t() @ a.double()) < 1e-5


@pytest.mark.parametrize("n,act,xw", [(128, "prelu", 128), (64, "prelu", 128), (64, "relu", 128), (128, "none", 128),
                                      (64, "prelu", 64), (64, "none", 256)])
def test_dw_tc_activation_on_load(n, act, xw):
    """X = act(pre) (layers 1 and 2 of the edge MLP; n = 64 against 128 columns runs with the operand roles
    swapped, n = 64 otherwise as a zero-padded 128-row operand)."""
    n_rows = 10_007
    g = torch.Generator(device="cpu").manual_seed(5)
    delta = torch.randn(n_rows, n, generator=g).to(DEV)
    pre = torch.randn(n_rows, xw, generator=g).to(DEV)
    slope = torch.tensor([0.3], device=DEV)
    code = _act_code(act)
    rows = lib.make_rows([(pre, None, xw, xw, code, slope if act == "prelu" else None)])
    grid = lib.mlp_layer_bwd_dw_tc_grid(n_rows)
    part = torch.full((grid * n * xw,), float("nan"), device=DEV)
    lib.mlp_layer_bwd_dw_tc(delta, n_rows, n, rows, part)
    torch.cuda.synchronize()
    (b,) = _reduce(part, grid, n, [xw])
    x = pre.double()
    x = {"prelu": torch.where(x > 0, x, 0.3 * x), "relu": x.clamp_min(0), "none": x}[act]
    assert _rel(b, delta.double().t() @ x) < 1e-5
    # and the deterministic device-side reduction the training step uses
    gw = torch.zeros(n, xw, device=DEV)
    lib.reduce_partials(part, grid, n * xw, 0, n * xw, xw, xw, gw, xw, 0)
    torch.cuda.synchronize()
    assert _rel(gw, delta.double().t() @ x) < 1e-5


def test_dw_tc_is_deterministic():
    n_rows = 50_000
    delta = torch.randn(n_rows, 128, device=DEV)
    pre = torch.randn(n_rows, 128, device=DEV)
    rows = lib.make_rows([(pre, None, 128, 128, 0, None)])
    grid = lib.mlp_layer_bwd_dw_tc_grid(n_rows)
    outs = []
    for _ in range(2):
        part = torch.empty(grid * 128 * 128, device=DEV)
        lib.mlp_layer_bwd_dw_tc(delta, n_rows, 128, rows, part)
        torch.cuda.synchronize()
        outs.append(part.clone())
    assert torch.equal(outs[0], outs[1])


def test_tc_rejects_unsupported_shapes():
    delta = torch.randn(10, 96, device=DEV)
    rows = lib.make_rows([(delta, None, 96, 96, 0, None)])
    part = torch.empty(96 * 96, device=DEV)
    with pytest.raises(RuntimeError):
        lib.mlp_layer_bwd_dw_tc(delta, 10, 96, rows, part)
    with pytest.raises(RuntimeError):
        lib.mlp_layer_bwd_dx_tc(delta, 10, 96, delta, 96, 0, 64, 64, part, False)


@pytest.mark.parametrize("act", ["prelu", "relu", "none"])
@pytest.mark.parametrize("n_rows,n,ko,split", [(1, 128, 128, 128), (5000, 64, 64, 64), (70_001, 128, 128, 64), (33_333, 64, 128, 128)])
def test_dx_tc_fused_delta_matches_separate_pass(n_rows, n, ko, split, act):
    """dx_tc forming delta = dh ⊙ act'(pre) itself: delta bit-identical to the element-wise CUDA-core pass, dx equal to
    the unfused tensor-core result, bias / slope partial sums within fp32 summation-order noise."""
    g = torch.Generator(device="cpu").manual_seed(n_rows + n)
    dh0 = torch.randn(n_rows, n, generator=g).to(DEV)
    pre = torch.randn(n_rows, n, generator=g).to(DEV)
    w = (torch.randn(n, 256, generator=g) * 0.1).to(DEV)
    slope = torch.tensor([0.2], device=DEV)
    code = _act_code(act)
    sl = slope if act == "prelu" else None
    # reference: separate delta pass + unfused dx_tc
    dh_a = dh0.clone()
    grid_a = lib.mlp_layer_bwd_dx_grid(n_rows)
    part_a = torch.zeros(grid_a * (n + 1), device=DEV)
    lib.mlp_layer_bwd_dx(dh_a, pre, code, sl, n_rows, n, w, 256, 0, 16, 16, None, False, True, part_a)
    outs_a = [torch.zeros(n_rows, split, device=DEV), torch.ones(n_rows, ko - split, device=DEV) if split != ko else None]
    lib.mlp_layer_bwd_dx_tc(dh_a, n_rows, n, w, 256, 32, ko, ko, outs_a[0], False, outs_a[1], True, split)
    # fused
    dh_b = dh0.clone()
    grid_b = lib.mlp_layer_bwd_dx_tc_grid(n_rows)
    part_b = torch.full((grid_b * (n + 1),), float("nan"), device=DEV)
    outs_b = [torch.zeros(n_rows, split, device=DEV), torch.ones(n_rows, ko - split, device=DEV) if split != ko else None]
    assert lib.mlp_layer_bwd_dx_tc_fused(dh_b, pre, code, sl, n_rows, n, w, 256, 32, ko, ko, outs_b[0], False, outs_b[1], True,
                                         split, part_b) == grid_b
    torch.cuda.synchronize()
    assert torch.equal(dh_a, dh_b)
    assert torch.equal(outs_a[0], outs_b[0])
    if outs_a[1] is not None:
        assert torch.equal(outs_a[1], outs_b[1])
    sa = part_a.view(grid_a, n + 1).double().sum(0)
    sb = part_b.view(grid_b, n + 1).double().sum(0)
    ref = dh_a.double().sum(0)
    assert float((sb[:n] - ref).abs().max()) <= 1e-5 * float(dh_a.double().abs().sum(0).max()) + 1e-12
    assert float((sa[:n] - sb[:n]).abs().max()) <= 1e-5 * float(dh_a.double().abs().sum(0).max()) + 1e-12
    if act == "prelu":
        ref_s = (dh0.double() * pre.double())[pre <= 0].sum()
        assert abs(float(sb[n] - ref_s)) <= 1e-5 * float((dh0.double() * pre.double()).abs().sum()) + 1e-12


def _guarded(n_elems, fill=float("nan")):
    """A tensor of n_elems floats embedded between two guard bands; returns (view, check) — check() asserts that the
    guards are untouched (the pool has no compute-sanitizer: out-of-range writes are caught this way)."""
    pad = 4096
    buf = torch.full((n_elems + 2 * pad,), 12345.0, device=DEV)
    view = buf[pad:pad + n_elems]
    view.fill_(fill)

    def check():
        torch.cuda.synchronize()
        assert bool((buf[:pad] == 12345.0).all()) and bool((buf[pad + n_elems:] == 12345.0).all()), "guard band overwritten"
    return view, check


@pytest.mark.parametrize("n_rows", [1, 127, 129, 4097, 33_001])
def test_tensor_core_training_kernels_stay_inside_their_buffers(n_rows):
    g = torch.Generator(device="cpu").manual_seed(n_rows)
    n = 128
    dh = torch.randn(n_rows, n, generator=g).to(DEV)
    pre = torch.randn(n_rows, n, generator=g).to(DEV)
    w = (torch.randn(n, 320, generator=g) * 0.1).to(DEV)
    slope = torch.tensor([0.2], device=DEV)
    checks = []
    # dx (two 64-wide blocks) with the fused delta pass
    dx0, c = _guarded(n_rows * 64); checks.append(c)
    dx1, c = _guarded(n_rows * 64); checks.append(c)
    grid = lib.mlp_layer_bwd_dx_tc_grid(n_rows)
    part, c = _guarded(grid * (n + 1)); checks.append(c)
    delta, c = _guarded(n_rows * n); checks.append(c)
    delta.copy_(dh.flatten())
    lib.mlp_layer_bwd_dx_tc_fused(delta, pre, lib.ACT_CODES["prelu"], slope, n_rows, n, w, 320, 64, 128, 128, dx0, False, dx1,
                                  False, 64, part)
    # dW with gathered segments
    n_nodes = max(n_rows // 3, 2)
    xs = torch.randn(n_nodes, 64, generator=g).to(DEV)
    src = torch.randint(0, n_nodes, (n_rows,), generator=g).to(torch.int32).to(DEV)
    rows = lib.make_rows([(xs, src, 64, 64, 0, None), (xs, src, 64, 64, 0, None), (pre, None, 128, 128, 1, slope)])
    gdw = lib.mlp_layer_bwd_dw_tc_grid(n_rows)
    pdw, c = _guarded(gdw * n * 256); checks.append(c)
    assert lib.mlp_layer_bwd_dw_tc(delta, n_rows, n, rows, pdw) == gdw
    for c in checks:
        c()
    assert bool(torch.isfinite(dx0).all()) and bool(torch.isfinite(dx1).all()) and bool(torch.isfinite(pdw).all())


def test_empty_inputs_are_no_ops():
    """Zero rows / zero edges: every tensor-core training entry point returns without touching its outputs."""
    z = torch.zeros(1, 128, device=DEV)
    w = torch.zeros(128, 128, device=DEV)
    out = torch.full((4, 128), 3.0, device=DEV)
    part = torch.full((1024,), 3.0, device=DEV)
    lib.mlp_layer_bwd_dx_tc(z, 0, 128, w, 128, 0, 128, 128, out, False)
    assert lib.mlp_layer_bwd_dx_tc_fused(z, z, 0, None, 0, 128, w, 128, 0, 128, 128, out, False, None, False, None, part) == 0
    rows = lib.make_rows([(z, None, 128, 128, 0, None)])
    assert lib.mlp_layer_bwd_dw_tc(z, 0, 128, rows, part) == 0
    torch.cuda.synchronize()
    assert bool((out == 3.0).all()) and bool((part == 3.0).all())
