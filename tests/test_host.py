"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import os
import re

import numpy as np
import pytest
import torch

import mswe_gnn_b200  # noqa: F401
from helpers import REF_CONFIG_MODELS
from mswe_gnn_b200 import lib
from mswe_gnn_b200.utils.data import Batch
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh, tri_level_sizes
from oracle import plan_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "swe_gnn_b200.h")).read()
    declared = set(re.findall(r"\b(swe_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.SIGNATURES), (declared ^ set(lib.SIGNATURES))
    assert l.swe_abi_version() == 1 and l.swe_build_arch() == b"sm_100a"


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.load(str(tmp_path / "nope.so"))


def test_models_refuse_cpu_tensors():
    from mswe_gnn_b200.models.gnn import MSGNN
    m = MSGNN(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **{**REF_CONFIG_MODELS, "hid_features": 16})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            m(make_tri_mesh(8, 4, 3))


def test_unsupported_options_raise():
    from mswe_gnn_b200.models.gnn import GNN, MSGNN
    from mswe_gnn_b200.models.models import activation_functions, make_mlp
    with pytest.raises(NotImplementedError):
        MSGNN(8, 1, 3, learned_pooling=True, previous_t=3)
    with pytest.raises(NotImplementedError):
        GNN(8, 1, type_GNN="GAT", previous_t=3)
    with pytest.raises(ValueError):
        GNN(8, 1, type_GNN="nonsense", previous_t=3)
    with pytest.raises(NotImplementedError):
        make_mlp(4, 4, layer_norm=True)
    with pytest.raises(AttributeError):
        activation_functions("gelu")
    with pytest.raises(AssertionError):
        MSGNN(8, 1, 3, K=[1, 2], previous_t=3)


def test_synthetic_mesh_sizes_match_survey():
    d = make_tri_mesh(32, 24, 4)
    assert d.x.shape[0] == 2044 and d.edge_index.shape[1] == 5914
    assert d.node_ptr.tolist() == [0, 1537, 1922, 2019, 2044]
    assert [s[0] for s in tri_level_sizes(712, 712, 4)] == [1013889, 253473, 63369, 15843]
    assert sum(s[1] for s in tri_level_sizes(712, 712, 4)) == 4034374
    assert sum(s[0] for s in tri_level_sizes(2832, 2832, 4)) == 21303724
    ei = d.edge_index[:, :4496]
    fwd = set(map(tuple, ei.t().tolist()))
    assert all((b, a) in fwd for a, b in fwd)                       # undirected part is symmetric
    assert d.edge_index[:, 4496].tolist() == [1536, 0]              # ghost -> face 0
    key = ei[0] * 10000 + ei[1]
    assert bool((key[1:] > key[:-1]).all())                         # (row, col) sorted
    c, f = d.intra_mesh_edge_index[:, :1536]
    assert bool((c[1:] >= c[:-1]).all()) and torch.bincount(c - 1537).tolist() == [4] * 384


def test_plan_oracle_loops_vs_numpy():
    rng = np.random.default_rng(0)
    row = rng.integers(0, 40, 300)
    col = rng.integers(10, 30, 300)
    a = P.stable_dst_csr(row, col, None, 10, 20)
    b = P.stable_dst_csr_loops(row.tolist(), col.tolist(), 10, 20)
    for x, y in zip(a, b):
        assert x.tolist() == list(y)
    # stable order == scatter_add_ order: sequential in-segment sums reproduce it bit for bit
    vals = torch.randn(300, 4)
    ref = torch.zeros(20, 4).scatter_add_(0, torch.from_numpy(col - 10).view(-1, 1).expand(-1, 4), vals)
    rowptr, _, _, eid = a
    out = torch.zeros(20, 4)
    for i in range(20):
        acc = torch.zeros(4)
        for p in range(rowptr[i], rowptr[i + 1]):
            acc = acc + vals[eid[p]]
        out[i] = acc
    assert torch.equal(out, ref)


def test_batch_collation_and_permutation():
    graphs = [make_tri_mesh(8, 8, 3, seed=s) for s in range(3)]
    b = Batch.from_data_list(graphs)
    n = [g.x.shape[0] for g in graphs]
    assert b.ptr.tolist() == [0, n[0], n[0] + n[1], sum(n)]
    assert torch.equal(b.edge_index[:, graphs[0].edge_index.shape[1]:][:, :5], graphs[1].edge_index[:, :5] + n[0])
    assert torch.equal(b.node_ptr[:4], graphs[0].node_ptr)          # *_ptr are NOT shifted by collation
    from mswe_gnn_b200.training.train import adapt_batch_training
    t = adapt_batch_training(b)
    assert t.node_ptr.shape == (3, 4) and int(t.node_ptr[-1, -1]) == sum(n)
    perm, inv = P.batch_permutation(t.node_ptr.numpy())
    assert sorted(perm.tolist()) == list(range(sum(n)))
    s0 = graphs[0].node_ptr[1].item()
    assert perm[:s0].tolist() == list(range(s0)) and perm[s0] == n[0]        # scale-major, graph-minor
    assert t.node_BC.tolist() == [s0 - 1, n[0] + s0 - 1, n[0] + n[1] + s0 - 1]


def test_adapted_batch_is_cached_per_topology_and_follows_the_values():
    """training_step adapts a collated batch once per topology (the reference re-adapts every step); x / y / BC of the
    current batch are used, and another collated batch gets its own adaptation."""
    from mswe_gnn_b200.training.train import _adapt_cached, adapt_batch_training
    graphs = [make_tri_mesh(8, 8, 3, rollout_steps=2, seed=s) for s in range(3)]
    b = Batch.from_data_list(graphs)
    t1 = _adapt_cached(b)
    b.x = b.x + 1.0                                                   # new values, same topology tensors
    t2 = _adapt_cached(b)
    assert t2.edge_index is t1.edge_index and t2.node_ptr is t1.node_ptr          # adapted topology reused
    assert t2.x is b.x and t1.x is not b.x                                          # values follow the batch
    ref = adapt_batch_training(b)
    for name in ("edge_index", "edge_attr", "node_ptr", "edge_ptr", "intra_mesh_edge_index", "intra_edge_ptr", "node_BC"):
        assert torch.equal(getattr(t2, name), getattr(ref, name)), name
    fin = t2._finest_rows
    rows = torch.zeros_like(fin)
    for r in ref.node_ptr.tolist():
        rows[r[0]:r[1]] = True
    assert torch.equal(fin, rows)
    b2 = Batch.from_data_list([make_tri_mesh(8, 8, 3, rollout_steps=2, seed=7), make_tri_mesh(8, 8, 3, rollout_steps=2, seed=8)])
    t3 = _adapt_cached(b2)
    assert t3.edge_index is not t1.edge_index and int(t3.node_ptr.shape[0]) == 2


@pytest.mark.parametrize("type_loss", ["RMSE", "MAE"])
def test_sync_free_loss_equals_boolean_indexed_loss(type_loss):
    """The mask-weighted form of the loss (no data-dependent shapes, no host reads) against the reference's boolean
    indexing (training/loss.py:49-118), multiscale batch and single graph, values and gradients."""
    from mswe_gnn_b200.training.loss import loss_function
    from mswe_gnn_b200.training.train import _adapt_cached
    torch.manual_seed(0)
    b = Batch.from_data_list([make_tri_mesh(8, 8, 3, rollout_steps=1, seed=s) for s in range(3)])
    t = _adapt_cached(b)
    n = t.x.shape[0]
    real = torch.rand(n, 2) * (torch.rand(n, 1) < 0.6)
    out = {}
    for cached in (True, False):
        preds = (real + torch.randn(n, 2, generator=torch.Generator().manual_seed(1)) * (torch.rand(n, 1, generator=torch.Generator().manual_seed(2)) < 0.7)).requires_grad_(True)
        data = t
        if not cached:
            data = t.__class__.__new__(t.__class__)
            data.__dict__.update({k: v for k, v in t.__dict__.items() if k != "_finest_rows"})
        loss = loss_function(preds, real, data, None, type_loss=type_loss, only_where_water=True, velocity_scaler=7.0)
        loss.backward()
        out[cached] = (loss.detach(), preds.grad.clone())
    assert torch.allclose(out[True][0], out[False][0], rtol=1e-6, atol=1e-8)
    assert torch.allclose(out[True][1], out[False][1], rtol=1e-5, atol=1e-9)


def test_adapt_batch_device_equals_adapt_batch_training():
    """The vectorised batch adaptation (one gather per edge list) against the mirror of the reference's Python slicing,
    field by field, bit for bit — multi-scale batches with ragged graphs and a single-scale batch."""
    import torch
    from mswe_gnn_b200.training.train import adapt_batch_device, adapt_batch_training
    from mswe_gnn_b200.utils.data import Batch
    from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh
    cases = [[make_tri_mesh(8 * (1 + s % 2), 8, 3, seed=s, orphan_every=5 if s == 1 else 0, extra_parent_every=4 if s == 2 else 0)
              for s in range(4)],
             [make_tri_mesh(16, 8, 4, seed=7)],
             [make_single_scale_mesh(6 + s, 5, seed=s) for s in range(3)]]
    for graphs in cases:
        batch = Batch.from_data_list(graphs)
        a, b = adapt_batch_training(batch), adapt_batch_device(batch)
        assert sorted(a.keys()) == sorted(b.keys())
        for k in a.keys():
            va, vb = getattr(a, k), getattr(b, k)
            if torch.is_tensor(va):
                assert va.shape == vb.shape and torch.equal(va.to(torch.int64) if not va.is_floating_point() else va,
                                                            vb.to(torch.int64) if not vb.is_floating_point() else vb), k
            elif isinstance(va, (int, float)):
                assert va == vb, k
