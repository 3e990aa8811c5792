"""tcgen05 row-MLP kernels (encoders, filter_matrix[0], decoder head; F = 64) against the exact-fp32 CUDA-core
kernels on identical inputs.  The 64->64 layers are 3xTF32 (swe_row_mlp_tc) or fp16 hi/lo with per-row scaling
(swe_row_mlp_tc16, swe_row_linear_tc16), fp32 accumulation: per layer ~1e-6 relative."""
import os

import pytest
import torch

import mswe_gnn_b200  # noqa: F401
from helpers import REF_CONFIG_MODELS, assert_close_masked, rel_l2
from mswe_gnn_b200 import lib
from mswe_gnn_b200.models.gnn import GNN, MSGNN
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _with_backend(name, fn):
    old = os.environ.get("MSWE_ROWMLP")
    os.environ["MSWE_ROWMLP"] = name
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("MSWE_ROWMLP", None)
        else:
            os.environ["MSWE_ROWMLP"] = old


@pytest.mark.parametrize("rm16", ["enc", "1", "0"])                # which two-layer stacks run on swe_row_mlp_tc16 (engine.rowmlp16_backend)
@pytest.mark.parametrize("nx,ny", [(16, 8), (320, 320)])          # the second: ~21 tiles per CTA in every pipeline
def test_msgnn_forward_tc_row_mlps_match_ffma(nx, ny, rm16, monkeypatch):
    monkeypatch.setenv("MSWE_ROWMLP16", rm16)
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **REF_CONFIG_MODELS)
    m = MSGNN(**ctor).to(DEV)
    d = make_tri_mesh(nx, ny, 4, seed=3).to(DEV)
    with torch.no_grad():
        a = _with_backend("tc", lambda: m(d))
        m._edge_cache = None
        b = _with_backend("ffma", lambda: m(d))
        m._edge_cache = None                                         # (the encoded edge features are cached per backend run)
        a2 = _with_backend("tc", lambda: m(d))
        for _ in range(4 if nx >= 320 else 1):                       # the pipelines are timing-sensitive: several repeats
            m._edge_cache = None
            assert torch.equal(a, _with_backend("tc", lambda: m(d)))
    assert torch.equal(a, a2)                                        # deterministic
    assert rel_l2(a, b) < 2e-5, rel_l2(a, b)
    assert_close_masked(a, b, 2e-4, 2e-5, "tc vs ffma row MLPs")


def test_encoders_and_w0_individually():
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=2, previous_t=3, **REF_CONFIG_MODELS)
    m = MSGNN(**ctor).to(DEV)
    d = make_tri_mesh(40, 24, 2, seed=5).to(DEV)
    d.x[torch.rand(d.x.shape[0], device=DEV) < 0.5, 2:] = 0         # dry nodes: the bias-free dynamic encoder must give exact 0
    plan = m._plans.get(d, 2, True)
    N, n0 = plan.n_nodes, plan.scale_n[0]
    outs = {}
    for be in ("tc", "ffma"):
        xs = torch.zeros(N, 64, device=DEV); xd = torch.zeros(N, 64, device=DEV)
        _with_backend(be, lambda: m._encode_nodes(d.x.contiguous(), plan, n0, xs, xd))
        outs[be] = (xs, xd)
    for t_tc, t_ff in zip(outs["tc"], outs["ffma"]):
        scale = float(t_ff.abs().max())
        assert float((t_tc - t_ff).abs().max()) <= 4e-6 * scale
    dry = (d.x[:n0, 2:] == 0).all(1)
    assert float(outs["tc"][1][:n0][dry].abs().max()) == 0.0
    # filter_matrix[0]
    la = m.gnn_processor[0].launcher()
    x = torch.randn(N, 64, device=DEV)
    o_tc = torch.zeros(N, 64, device=DEV); o_ff = torch.zeros(N, 64, device=DEV)
    la.w0_tc.linear(x, 5, N - 9, o_tc)
    lib.node_linear_fwd(x, 5, N - 9, la.filters.tensors()[0], o_ff, 64)
    assert float((o_tc - o_ff).abs().max()) <= 4e-6 * float(o_ff.abs().max())
    assert float(o_tc[:5].abs().max()) == 0 and float(o_tc[N - 4:].abs().max()) == 0


def test_gnn_rollout_tc_vs_ffma_head_and_window_shift():
    from mswe_gnn_b200.training.train import rollout_test
    gc = {k: v for k, v in REF_CONFIG_MODELS.items() if k not in ("learned_pooling", "skip_connections")}
    m = GNN(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=2, **{**gc, "learned_residuals": "all"}).to(DEV)
    d = make_single_scale_mesh(30, 20, rollout_steps=4, seed=2).to(DEV)
    a = _with_backend("tc", lambda: rollout_test(m, d))
    m._edge_cache = None
    b = _with_backend("ffma", lambda: rollout_test(m, d))
    assert rel_l2(a, b) < 1e-4, rel_l2(a, b)
