"""CPU, only where /root/reference exists: the oracle against the REAL reference imported under
stubs (oracle/ref_stubs.py) — forward, rollout helpers, batch adaptation, loss and gradients."""
import copy

import pytest
import torch
import yaml

from oracle import ref_stubs
from oracle import swe_oracle as O

pytestmark = pytest.mark.skipif(not ref_stubs.reference_available(), reason="/root/reference not present")

import mswe_gnn_b200  # noqa: E402,F401
from helpers import REF_CONFIG_MODELS  # noqa: E402
from mswe_gnn_b200.utils.data import Batch  # noqa: E402
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh  # noqa: E402


@pytest.fixture(scope="module")
def R():
    return ref_stubs.load_reference()


def test_config_copy_matches_reference_yaml():
    cfg = yaml.safe_load(open(ref_stubs.REFERENCE_ROOT + "/config.yaml"))["models"]
    cfg.pop("model_type")
    assert cfg == REF_CONFIG_MODELS


@pytest.mark.parametrize("kw", [dict(), dict(with_WL=False, learned_residuals=False, K=[1, 2, 3]),
                                dict(skip_connections=False, gnn_activation=None, mlp_activation="relu", mlp_layers=1),
                                dict(with_gradient=False, normalize=False, learned_residuals=None)])
def test_msgnn_forward_bit_exact(R, kw):
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3,
                **{**REF_CONFIG_MODELS, "hid_features": 16, **kw})
    m = R.MSGNN(**ctor)
    spec = O.ModelSpec("MSGNN", **ctor)
    d = make_tri_mesh(16, 8, 3, seed=2, orphan_every=5, link_ghosts=True)
    with torch.no_grad():
        assert torch.equal(m(ref_stubs.to_stub(d)), O.forward(m.state_dict(), spec, d))
        assert torch.equal(m(ref_stubs.to_stub(d)), O.forward(m.state_dict(), spec, d, hoisted=True))


def test_gnn_forward_bit_exact(R):
    gc = {k: v for k, v in REF_CONFIG_MODELS.items() if k not in ("learned_pooling", "skip_connections")}
    ctor = dict(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=3, **{**gc, "hid_features": 32, "K": 2})
    m = R.GNN(**ctor)
    d = make_single_scale_mesh(12, 10, seed=4)
    with torch.no_grad():
        assert torch.equal(m(ref_stubs.to_stub(d)), O.forward(m.state_dict(), O.ModelSpec("GNN", **ctor), d))


def test_rollout_helpers_and_scale_mask(R):
    d = make_tri_mesh(16, 8, 3, rollout_steps=2)
    x1, x2 = d.x.clone(), d.x.clone()
    a = R.apply_boundary_condition(x1[:, -6:], d.BC[:, :, 0], d.node_BC, type_BC=2)
    b = O.apply_boundary_condition(x2[:, -6:], d.BC[:, :, 0], d.node_BC, 2)
    assert torch.equal(a, b) and torch.equal(x1, x2)
    p = torch.rand(d.x.shape[0], 2)
    assert torch.equal(R.use_prediction(x1, p, 3), O.use_prediction(x2, p, 3))
    assert torch.equal(R.create_scale_mask(d.x.shape[0], 3, d.node_ptr, ref_stubs.to_stub(d)),
                       O.scale_mask(d.x.shape[0], d.node_ptr))
    from mswe_gnn_b200.utils import dataset as our
    assert torch.equal(our.use_prediction(x1, p, 3), R.use_prediction(x1, p, 3))
    assert torch.equal(our.create_scale_mask(d.x.shape[0], 3, d.node_ptr), O.scale_mask(d.x.shape[0], d.node_ptr))
    with pytest.raises(ValueError):
        our.apply_boundary_condition(x1[:, -6:], d.BC[:, :, 0], d.node_BC, type_BC=3)
    with pytest.raises(ValueError):
        R.apply_boundary_condition(x1[:, -6:], d.BC[:, :, 0], d.node_BC, type_BC=3)


def test_batched_forward_matches_reference_layout(R):
    """Adapted multiscale batch (train.py:14-65): our adaptation == the oracle's == what the
    reference forward accepts; batched prediction == per-graph predictions."""
    from mswe_gnn_b200.training.train import adapt_batch_training
    graphs = [make_tri_mesh(16, 8, 3, seed=s) for s in (1, 2, 3)]
    batch = Batch.from_data_list(graphs)
    ours = adapt_batch_training(batch)
    orc = O.adapt_batch(batch, graphs)
    for k in ("node_ptr", "edge_index", "edge_attr", "edge_ptr", "intra_edge_ptr", "intra_mesh_edge_index", "node_BC"):
        assert torch.equal(getattr(ours, k), getattr(orc, k)), k
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3,
                **{**REF_CONFIG_MODELS, "hid_features": 16})
    m = R.MSGNN(**ctor)
    stub = ref_stubs.to_stub(ours, batch=True)
    with torch.no_grad():
        out_b = m(stub)
        assert torch.equal(out_b, O.forward(m.state_dict(), O.ModelSpec("MSGNN", **ctor), ours))
        for g, (lo, hi) in zip(graphs, zip(batch.ptr[:-1], batch.ptr[1:])):
            assert torch.allclose(out_b[lo:hi], m(ref_stubs.to_stub(g)), atol=1e-6)


def test_training_step_loss_and_grads(R):
    """Oracle training_step (BPTT over 2 steps + wet-cell RMSE) == the reference pieces wired as
    in LightningTrainer.training_step (train.py:125-145)."""
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3,
                **{**REF_CONFIG_MODELS, "hid_features": 16, "K": 2})
    m = R.MSGNN(**ctor)
    d = make_tri_mesh(16, 8, 3, rollout_steps=2, seed=6)
    t = ref_stubs.to_stub(d).clone()
    losses = []
    for i in range(2):
        t.x[:, -6:] = R.apply_boundary_condition(t.x[:, -6:], t.BC[:, :, i], t.node_BC, type_BC=t.type_BC)
        p = m(t)
        t.x = R.use_prediction(t.x, p, 3)
        losses.append(R.loss_function(p, t.y[:, :, i], t, t.BC[:, -2:, i + 1].mean(1), type_loss="RMSE",
                                      only_where_water=True, conservation=0, velocity_scaler=7))
    ref_loss = torch.stack(losses).mean()
    ref_loss.backward()
    ref_grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    loss = O.training_step(sd, O.ModelSpec("MSGNN", **ctor), d, 2, type_loss="RMSE", only_where_water=True,
                           velocity_scaler=7.0)
    loss.backward()
    assert torch.equal(loss.detach(), ref_loss.detach())
    for k, g in ref_grads.items():
        assert torch.allclose(sd[k].grad, g, rtol=1e-6, atol=1e-9), k
