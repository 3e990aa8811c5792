"""CPU: the oracle (oracle/swe_oracle.py) against the reference-generated golden vectors, the
published parameter counts and the constructor/RNG parity of our model classes."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, build_model, load_fixture, make_mesh, sd_sha, spec_of
from oracle import swe_oracle as O

MODEL_FIXTURES = ["msgnn_k4f64_cfg1", "msgnn_k213f16_irregular", "gnn_k3f64_single"]


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_constructor_reproduces_reference_weights(name):
    meta, z = load_fixture(name)
    m = build_model(meta)
    assert sd_sha(m.state_dict()) == meta["weights_sha"]            # same names, shapes, RNG order, values
    assert sum(p.numel() for p in m.parameters()) == meta["n_params"]


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_oracle_forward_and_rollout_bit_exact(name):
    meta, z = load_fixture(name)
    m = build_model(meta)
    sd, spec, d = m.state_dict(), spec_of(meta), make_mesh(meta)
    with torch.no_grad():
        out = O.forward(sd, spec, d)
    assert np.array_equal(out.numpy(), z["forward"])
    T = z["rollout"].shape[-1]
    roll = O.rollout(sd, spec, d, steps=T)
    assert np.array_equal(roll.numpy(), z["rollout"])
    # the hoisted / mask-free formulation used by the CUDA path is the same function
    with torch.no_grad():
        out_h = O.forward(sd, spec, d, hoisted=True)
    assert np.array_equal(out_h.numpy(), z["forward"])


def test_oracle_trained_checkpoint_rollout():
    meta, z = load_fixture("msgnn_k4f32_trained_drybed")
    m = build_model(meta, z)
    assert sd_sha(m.state_dict()) == meta["weights_sha"]
    assert meta["n_params"] == 203949                                # overview_MSGNN.csv:11
    roll = O.rollout(m.state_dict(), spec_of(meta), make_mesh(meta), steps=8)
    assert np.array_equal(roll.numpy(), z["rollout"])
    # the fixture is physically meaningful: the flood front advances from the inflow cell
    wet = (z["rollout"][:1537, 0, :] > 0).mean(0)
    assert wet[-1] > wet[0] > 0


def test_oracle_swegnn_operator_variants():
    meta, z = load_fixture("swegnn_operator")
    from mswe_gnn_b200.models.gnn import SWEGNN
    xs, xd, ea = (torch.from_numpy(z[k]) for k in ("x_s", "x_d", "edge_attr"))
    ei = torch.from_numpy(z["edge_index"])
    for i, v in enumerate(meta["variants"]):
        kw = v["kw"]
        torch.manual_seed(v["seed"])
        op = SWEGNN(16, 16, n_layers=2, activation="prelu", bias=True, **kw)
        assert sd_sha(op.state_dict()) == v["weights_sha"]
        sd = {"op." + k: t for k, t in op.state_dict().items()}
        out = O.swegnn(sd, "op", xs, xd, ei, ea if kw["edge_features"] else None, kw["K"], 2, "prelu",
                       kw["edge_features"], normalize=kw.get("normalize", True),
                       with_filter_matrix=kw.get("with_filter_matrix", True),
                       with_gradient=kw.get("with_gradient", True), upwind_mode=kw.get("upwind_mode", False))
        assert np.array_equal(out.numpy(), z[f"out{i}"]), f"variant {i}"


def test_published_parameter_counts():
    """results/Pareto_front/overview_MSGNN.csv / overview_GNN.csv 'total parameters' column."""
    from mswe_gnn_b200.models.gnn import GNN, MSGNN
    from helpers import REF_CONFIG_MODELS as C
    counts = json.load(open(os.path.join(GOLDEN, "param_counts.json")))
    for key, n in counts["MSGNN"].items():
        K, F = map(int, key.split(","))
        m = MSGNN(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **{**C, "K": K, "hid_features": F})
        assert sum(p.numel() for p in m.parameters()) == n, key
    gc = {k: v for k, v in C.items() if k not in ("learned_pooling", "skip_connections")}
    for key, n in counts["GNN"].items():
        F, K = map(int, key.split(","))
        m = GNN(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=2, **{**gc, "K": K, "hid_features": F})
        assert sum(p.numel() for p in m.parameters()) == n, key


def test_fp32_vs_fp64_drift_yardstick():
    """The tolerance yard-stick of SURVEY §8c: the reference's own fp32 path drifts from its fp64
    path over a rollout; the GPU tests bound our drift by a multiple of this."""
    meta, z = load_fixture("msgnn_k4f32_trained_drybed")
    r32, r64 = z["rollout"].astype(np.float64), z["rollout_fp64"]
    d0 = np.linalg.norm(r32[..., 0] - r64[..., 0]) / np.linalg.norm(r64[..., 0])
    d7 = np.linalg.norm(r32[..., 7] - r64[..., 7]) / np.linalg.norm(r64[..., 7])
    assert d0 < 5e-6 and d7 < 1e-3
