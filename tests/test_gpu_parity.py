"""GPU: the CUDA path (through the C ABI) against the oracle and the reference-generated golden
vectors.  Tolerances (fp32, stated per north_star): one forward = ~40 fused layers, each within
rel 1e-5 of fp32 arithmetic in a different summation order → forward rel-L2 <= 2e-5 and
element-wise |Δ| <= 2e-5 + 1e-4·|ref| except at wet/dry threshold flips; rollout drift is bounded
by a multiple of the reference's own fp32-vs-fp64 drift."""
import numpy as np
import pytest
import torch

from helpers import (REF_CONFIG_MODELS, assert_close_masked, build_model, load_fixture, make_mesh, rel_l2, spec_of)
from mswe_gnn_b200 import lib
from mswe_gnn_b200.utils.data import Batch
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh
from oracle import plan_oracle as P
from oracle import swe_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FWD_RTOL, FWD_ATOL, FWD_L2 = 1e-4, 2e-5, 2e-5


# ------------------------------------------------------------------------------------------------
# integer work: bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed,n,e", [(0, 50, 400), (1, 1, 7), (2, 1000, 0), (3, 5000, 40000)])
def test_csr_build_bit_exact(seed, n, e):
    rng = np.random.default_rng(seed)
    row = rng.integers(0, n + 30, e)
    col = rng.integers(20, 20 + n, e)
    ref = P.stable_dst_csr(row, col, None, 20, n)
    got = lib.csr_build(torch.from_numpy(row).to(DEV), torch.from_numpy(col).to(DEV), None, 20, n, 0, n + 30)
    for a, b in zip(got, ref):
        assert np.array_equal(a.cpu().numpy(), b)
    # transposed (by row) with a node map
    nm = rng.permutation(n + 50).astype(np.int32)
    r2, c2 = rng.integers(0, n + 50, e), rng.integers(0, n + 50, e)
    ref = P.stable_dst_csr(r2, c2, nm, 0, n + 50, by_row=True)
    got = lib.csr_build(torch.from_numpy(r2).to(DEV), torch.from_numpy(c2).to(DEV), torch.from_numpy(nm).to(DEV),
                        0, n + 50, 0, n + 50, by_row=True)
    for a, b in zip(got, ref):
        assert np.array_equal(a.cpu().numpy(), b)


def test_csr_build_rejects_out_of_range_edges():
    row = torch.tensor([0, 1, 2], device=DEV)
    col = torch.tensor([1, 99, 0], device=DEV)
    with pytest.raises(ValueError, match="outside the node range"):
        lib.csr_build(row, col, None, 0, 3, 0, 3)


def test_plan_of_adapted_batch_bit_exact():
    from mswe_gnn_b200.plan import build_plan
    from mswe_gnn_b200.training.train import adapt_batch_training
    graphs = [make_tri_mesh(8, 8, 3, seed=s, orphan_every=5) for s in range(3)]
    t = adapt_batch_training(Batch.from_data_list(graphs)).to(DEV)
    plan = build_plan(t, 3, True)
    perm, inv = P.batch_permutation(t.node_ptr.cpu().numpy())
    assert np.array_equal(plan.perm.cpu().numpy(), perm) and np.array_equal(plan.inv.cpu().numpy(), inv)
    ep = t.edge_ptr.tolist()
    for s in range(3):
        ei = t.edge_index[:, ep[s]:ep[s + 1]].cpu().numpy()
        ref = P.stable_dst_csr(ei[0], ei[1], inv, plan.scale_lo[s], plan.scale_n[s])
        es = plan.edges[s]
        for a, b in zip((es.rowptr, es.src, es.dst, es.eid), ref):
            assert np.array_equal(a.cpu().numpy(), b)
    ip = t.intra_edge_ptr.tolist()
    for j in range(2):
        ie = t.intra_mesh_edge_index[:, ip[j]:ip[j + 1]].cpu().numpy()
        ref = P.stable_dst_csr(ie[1], ie[0], inv, plan.scale_lo[j + 1], plan.scale_n[j + 1])     # pool: by coarse
        for a, b in zip((plan.pool[j].rowptr, plan.pool[j].src, plan.pool[j].eid), (ref[0], ref[1], ref[3])):
            assert np.array_equal(a.cpu().numpy(), b)
        ref = P.stable_dst_csr(ie[0], ie[1], inv, plan.scale_lo[j], plan.scale_n[j])             # un-pool: by fine
        for a, b in zip((plan.unpool[j].rowptr, plan.unpool[j].src, plan.unpool[j].eid), (ref[0], ref[1], ref[3])):
            assert np.array_equal(a.cpu().numpy(), b)


def test_bad_topology_raises():
    d = make_tri_mesh(8, 8, 3).to(DEV)
    m = build_model(dict(model="MSGNN", ctor=dict(num_node_features=8, num_edge_features=1, num_scales=3,
                                                  previous_t=3, hid_features=16)), device=DEV)
    d.edge_index = d.edge_index.clone()
    d.edge_index[1, 0] = d.node_ptr[1]          # a scale-0 edge pointing into scale 1
    with pytest.raises(ValueError, match="same scale"):
        with torch.no_grad():
            m(d)


# ------------------------------------------------------------------------------------------------
# operator level
# ------------------------------------------------------------------------------------------------
def test_swegnn_operator_variants_vs_golden():
    from mswe_gnn_b200.models.gnn import SWEGNN
    meta, z = load_fixture("swegnn_operator")
    xs, xd, ea = (torch.from_numpy(z[k]).to(DEV) for k in ("x_s", "x_d", "edge_attr"))
    ei = torch.from_numpy(z["edge_index"]).to(DEV)
    for i, v in enumerate(meta["variants"]):
        kw = v["kw"]
        torch.manual_seed(v["seed"])
        op = SWEGNN(16, 16, n_layers=2, activation="prelu", bias=True, **kw).to(DEV)
        with torch.no_grad():
            out = op(xs, xd, ei, ea if kw["edge_features"] else None)
        ref = torch.from_numpy(z[f"out{i}"])
        assert rel_l2(out.cpu(), ref) < 1e-5, (i, rel_l2(out.cpu(), ref))
        assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-5), i


def test_hop_aggregation_is_bit_exact_given_identical_inputs():
    """propagate_hop without filter matrix performs exactly the reference's fp32 operations in
    the reference's order (mul, then sequential adds in edge order) -> bit-exact vs scatter_add_."""
    torch.manual_seed(0)
    d = make_single_scale_mesh(20, 12, seed=1)
    n, e = d.x.shape[0], d.edge_index.shape[1]
    o = torch.randn(n, 32)
    s = torch.randn(e, 32)
    row, col = d.edge_index
    ref_g = o + O.scatter_sum((o[col] - o[row]) * s, col, n)
    ref_n = o + O.scatter_sum(s * o[row], col, n)
    rowptr, src, dst, eid = lib.csr_build(row.to(DEV), col.to(DEV), None, 0, n, 0, n)
    s_csr = s.to(DEV)[eid.long()].contiguous()
    od = o.to(DEV)
    for with_grad, ref in ((1, ref_g), (0, ref_n)):
        out = torch.empty_like(od)
        lib.propagate_hop_fwd(od, od, s_csr, rowptr, src, 0, n, None, with_grad, 0, None, 0, None, out, 32)
        assert torch.equal(out.cpu(), ref)


def test_pool_mean_bit_exact():
    d = make_tri_mesh(16, 8, 2, orphan_every=3, extra_parent_every=4)
    n = d.x.shape[0]
    x = torch.randn(n, 16)
    coarse, fine = d.intra_mesh_edge_index
    ref = O.scatter_mean(x[fine], coarse, n)
    n0, n1 = int(d.node_ptr[1]), int(d.node_ptr[2] - d.node_ptr[1])
    rowptr, src, dst, eid = lib.csr_build(fine.to(DEV), coarse.to(DEV), None, n0, n1, 0, n0)
    out = torch.zeros(n, 16, device=DEV)
    lib.pool_mean_fwd(x.to(DEV), rowptr, src, n0, n1, out, 16)
    assert torch.equal(out.cpu()[n0:], ref[n0:])


# ------------------------------------------------------------------------------------------------
# model level
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["msgnn_k4f64_cfg1", "msgnn_k213f16_irregular", "gnn_k3f64_single"])
def test_forward_and_rollout_vs_golden(name):
    from mswe_gnn_b200.training.train import rollout_test
    meta, z = load_fixture(name)
    m = build_model(meta, z, DEV)
    d = make_mesh(meta).to(DEV)
    with torch.no_grad():
        out = m(d)
    ref = torch.from_numpy(z["forward"])
    assert rel_l2(out.cpu(), ref) < FWD_L2, rel_l2(out.cpu(), ref)
    assert_close_masked(out, ref, FWD_RTOL, FWD_ATOL, name + " forward")
    # input graph is not mutated by forward (gnn.py:269 clones)
    assert torch.equal(d.x.cpu(), make_mesh(meta).x)
    roll = rollout_test(m, d)
    ref_r = torch.from_numpy(z["rollout"])
    assert roll.shape == ref_r.shape
    for t in range(ref_r.shape[-1]):
        assert rel_l2(roll[..., t].cpu(), ref_r[..., t]) < FWD_L2 * (t + 1) * 4, (t, rel_l2(roll[..., t].cpu(), ref_r[..., t]))
    # eager loop == CUDA-graph replay, bit for bit (deterministic kernels)
    roll2 = rollout_test(m, d, use_cuda_graph=False)
    assert torch.equal(roll, roll2)


def test_trained_checkpoint_rollout_drift_bounded_by_fp64_yardstick():
    from mswe_gnn_b200.training.train import rollout_test
    meta, z = load_fixture("msgnn_k4f32_trained_drybed")
    m = build_model(meta, z, DEV)
    d = make_mesh(meta).to(DEV)
    roll = rollout_test(m, d).cpu().double().numpy()
    r32, r64 = z["rollout"].astype(np.float64), z["rollout_fp64"]
    for t in range(8):
        ours = np.linalg.norm(roll[..., t] - r64[..., t]) / np.linalg.norm(r64[..., t])
        yard = np.linalg.norm(r32[..., t] - r64[..., t]) / np.linalg.norm(r64[..., t])
        assert ours <= 4 * yard + 2e-6, (t, ours, yard)
    # flood front advances like the reference's
    wet_o, wet_r = (roll[:1537, 0, -1] > 0).mean(), (r32[:1537, 0, -1] > 0).mean()
    assert abs(wet_o - wet_r) < 5e-3


@pytest.mark.parametrize("kw", [dict(with_WL=False, learned_residuals=False, K=[1, 2, 3]),
                                dict(skip_connections=False, gnn_activation=None, mlp_activation="relu", mlp_layers=1),
                                dict(with_gradient=False, normalize=False, learned_residuals=None, mlp_layers=2),
                                dict(hid_features=32, mlp_activation="leakyrelu", gnn_activation="sigmoid"),
                                dict(hid_features=24, mlp_activation="elu", gnn_activation="swish", with_filter_matrix=False)])
def test_msgnn_option_matrix_vs_oracle(kw):
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3,
                **{**REF_CONFIG_MODELS, "hid_features": 16, **kw})
    meta = dict(model="MSGNN", ctor=ctor)
    m = build_model(meta, device=DEV)
    d = make_tri_mesh(16, 8, 3, seed=2, orphan_every=5, link_ghosts=True)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.forward(sd, spec_of(meta), d)
        out = m(d.to(DEV))
    assert rel_l2(out.cpu(), ref) < FWD_L2, rel_l2(out.cpu(), ref)
    assert_close_masked(out, ref, FWD_RTOL, FWD_ATOL, str(kw))


def test_batched_msgnn_vs_oracle_and_per_graph():
    from mswe_gnn_b200.training.train import adapt_batch_training, rollout_test
    graphs = [make_tri_mesh(16, 8, 3, seed=s, rollout_steps=2) for s in (1, 2, 3)]
    batch = Batch.from_data_list(graphs)
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **{**REF_CONFIG_MODELS, "hid_features": 16})
    meta = dict(model="MSGNN", ctor=ctor)
    m = build_model(meta, device=DEV)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    adapted = adapt_batch_training(batch)
    ref = O.rollout(sd, spec_of(meta), O.adapt_batch(batch, graphs), steps=2)
    roll = rollout_test(m, batch.to(DEV))
    assert rel_l2(roll.cpu(), ref) < 4 * FWD_L2, rel_l2(roll.cpu(), ref)
    with torch.no_grad():
        out_b = m(adapted.to(DEV)).cpu()
        for g, lo, hi in zip(graphs, batch.ptr[:-1].tolist(), batch.ptr[1:].tolist()):
            assert torch.equal(out_b[lo:hi], m(g.to(DEV)).cpu())      # deterministic: batching changes nothing


def test_batched_gnn_vs_oracle():
    gc = {k: v for k, v in REF_CONFIG_MODELS.items() if k not in ("learned_pooling", "skip_connections")}
    ctor = dict(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=2, **{**gc, "hid_features": 32, "K": 3})
    meta = dict(model="GNN", ctor=ctor)
    m = build_model(meta, device=DEV)
    graphs = [make_single_scale_mesh(12, 10, seed=s) for s in (1, 2)]
    batch = Batch.from_data_list(graphs)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.forward(sd, spec_of(meta), batch)
        out = m(batch.to(DEV))
    assert rel_l2(out.cpu(), ref) < FWD_L2
    assert_close_masked(out, ref, FWD_RTOL, FWD_ATOL, "batched GNN")


def test_empty_and_degenerate_inputs():
    """Zero wet nodes (all-dry bed without inflow) -> exact zeros, like the reference (the dynamic
    encoder has no bias, gnn.py:209-210); an isolated node and a node with in-degree 0 are fine."""
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **{**REF_CONFIG_MODELS, "hid_features": 16})
    meta = dict(model="MSGNN", ctor=ctor)
    m = build_model(meta, device=DEV)
    d = make_tri_mesh(8, 8, 3, wet="dry")
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.forward(sd, spec_of(meta), d)
        out = m(d.to(DEV)).cpu()
    assert torch.equal(out, ref) and float(out.abs().max()) == 0.0


def test_large_mesh_properties():
    """cfg3-scale mesh (1.35 M nodes): size-independent properties instead of an oracle run —
    determinism, finiteness, dry rows stay exactly zero-depth-masked, and the top-left 2k-node
    corner matches what the same weights give on... (locality: K·S hops only reach so far)."""
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **REF_CONFIG_MODELS)
    m = build_model(dict(model="MSGNN", ctor=ctor), device=DEV)
    d = make_tri_mesh(712, 712, 4, seed=0).to(DEV)
    with torch.no_grad():
        a = m(d)
        b = m(d)
    assert a.shape == (1346574, 2)
    assert bool(torch.isfinite(a).all())
    assert torch.equal(a, b), f"not deterministic: {int((a != b).sum())} entries differ"
    assert bool((a >= 0).all())
    h, q = a[:, 0], a[:, 1]
    assert bool(((h == 0) | (h.abs() > 1e-4)).all())                 # dry mask, models.py:79-91
    # linearity of the residual head under a dry bed: all-dry input -> all-zero output
    d.x[:, 2:] = 0
    with torch.no_grad():
        z = m(d)
    assert float(z.abs().max()) == 0.0


def test_rollout_test_reuses_runner_for_same_topology_with_new_values():
    """Repeated rollout_test calls on freshly loaded copies of the same mesh reuse plan / workspaces / captured step (cache
    keyed by a content hash of the topology): results equal a cold call bit for bit, also when node inputs, boundary
    series and edge attributes change between calls; earlier results are not overwritten."""
    from mswe_gnn_b200.training import train as T
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **REF_CONFIG_MODELS)
    m = build_model(dict(model="MSGNN", ctor=ctor), device=DEV)
    base = make_tri_mesh(24, 16, 3, seed=4, rollout_steps=5)
    variants = []
    for k in range(3):
        g = base.clone()
        if k:
            g.x[:, 2:] = g.x[:, 2:] * (1.0 + 0.3 * k)
            g.BC = g.BC * (1.0 + k)
            g.edge_attr = g.edge_attr * (1.0 - 0.2 * k)
        variants.append(g)
    T._RUNNER_CACHE.clear()
    warm = [T.rollout_test(m, g.to(DEV)) for g in variants]          # first call builds, the others re-bind
    assert len(T._RUNNER_CACHE) == 1
    cold = []
    for g in variants:
        T._RUNNER_CACHE.clear()
        cold.append(T.rollout_test(m, g.to(DEV)))
    for a, b in zip(warm, cold):
        assert torch.equal(a, b)
    assert not torch.equal(warm[0], warm[1])                         # the variants really differ, and the results are copies


def test_rollout_test_streams_predictions_to_pinned_host():
    """rollout_test(..., out_host=pinned [T, N, 2]) fills the host buffer step by step on a side stream: same bits as the
    returned device tensor, on the first call (eager step + capture) and on a cached-runner call; a wrong buffer is refused."""
    from mswe_gnn_b200.training.train import rollout_test
    from mswe_gnn_b200.utils.synthetic import make_tri_mesh
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **REF_CONFIG_MODELS)
    m = build_model(dict(model="MSGNN", ctor=ctor), device=DEV)
    d = make_tri_mesh(24, 20, 3, seed=4, rollout_steps=5).to(DEV)
    N, T = d.x.shape[0], d.y.shape[-1]
    host = torch.full((T, N, 2), float("nan")).pin_memory()
    for _ in range(2):
        host.fill_(float("nan"))
        pred = rollout_test(m, d, out_host=host)                       # [N, 2, T]
        torch.cuda.synchronize()
        assert torch.equal(host, pred.permute(2, 0, 1).cpu())
    with pytest.raises(ValueError):
        rollout_test(m, d, out_host=torch.empty(T, N, 2))              # not pinned
    with pytest.raises(ValueError):
        rollout_test(m, d, out_host=torch.empty(T + 1, N, 2).pin_memory())
