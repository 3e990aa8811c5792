"""GPU: the device-side remainder of the training step — loss + gradient kernels against the torch
mirror of the reference's loss.py, clip + AdamW on the flat buffers against torch.nn.utils.clip_grad_norm_ +
torch.optim.AdamW, and the whole step (eager and as one captured CUDA graph) against the oracle's training step followed
by a torch AdamW step."""
import copy

import pytest
import torch

from helpers import REF_CONFIG_MODELS, rel_l2
from mswe_gnn_b200 import lib
from mswe_gnn_b200.training.loss import loss_function
from mswe_gnn_b200.training.optim import FlatAdamW, device_loss
from mswe_gnn_b200.utils.data import Batch, Data
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh
from oracle import swe_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("type_loss", ["RMSE", "MAE"])
@pytest.mark.parametrize("only_where_water", [True, False])
@pytest.mark.parametrize("multiscale", [True, False])
def test_device_loss_matches_reference_loss_and_gradient(type_loss, only_where_water, multiscale):
    torch.manual_seed(0)
    n, T = 5000, 3
    pred = torch.rand(n, 2, device=DEV)
    y = torch.rand(n, 2, T, device=DEV)
    dry = torch.rand(n, device=DEV) < 0.4
    pred[dry] = 0
    y[dry] = 0                                                  # rows where pred == real: excluded by mask_on_water
    data = Data(x=torch.zeros(n, 8, device=DEV))
    rows = None
    if multiscale:
        data.node_ptr = torch.tensor([0, 3100, 4200, 5000])
        rows = torch.zeros(n, dtype=torch.bool, device=DEV)
        rows[:3100] = True
    p1 = pred.clone().requires_grad_(True)
    ref = loss_function(p1, y[:, :, 1], data, None, type_loss=type_loss, only_where_water=only_where_water, velocity_scaler=7.0)
    ref.backward()
    p2 = pred.clone().requires_grad_(True)
    ours = device_loss(p2, y[:, :, 1], rows, type_loss, only_where_water, 7.0)
    (2.5 * ours).backward()
    assert abs(float(ours) - float(ref)) <= 2e-6 * abs(float(ref))
    assert rel_l2(p2.grad / 2.5, p1.grad) < 2e-6
    if multiscale:
        assert float(p2.grad[3100:].abs().max()) == 0.0


def test_flat_adamw_matches_torch_clip_and_adamw():
    torch.manual_seed(1)
    net_a = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.PReLU(), torch.nn.Linear(64, 3)).to(DEV)
    net_b = copy.deepcopy(net_a)
    ref = torch.optim.AdamW(net_a.parameters(), lr=3e-3, weight_decay=0.01)
    ours = FlatAdamW(net_b, lr=3e-3, weight_decay=0.01, max_norm=1.0)
    for it in range(5):
        x = torch.randn(256, 40, device=DEV) * (10.0 if it % 2 == 0 else 0.01)      # clipped and un-clipped steps
        ref.zero_grad(set_to_none=True)
        net_a(x).square().sum().backward()
        n_ref = torch.nn.utils.clip_grad_norm_(net_a.parameters(), 1.0)
        ref.step()
        ours.zero_grad()
        net_b(x).square().sum().backward()
        ours.step()
        assert abs(float(ours.last_grad_norm) - float(n_ref)) <= 1e-5 * float(n_ref)
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            assert rel_l2(pb.detach(), pa.detach()) < 2e-6, it
        if it == 2:
            ours.set_lr(1e-3)
            for gparam in ref.param_groups:
                gparam["lr"] = 1e-3
    assert float(ours.state[0]) == 5.0


def _oracle_steps(model_type, ctor, data_or_batch, graphs, n_steps, lr):
    """n_steps of: oracle training_step (fp32 torch autograd on the reference formulation) + clip 1.0 + torch AdamW."""
    from mswe_gnn_b200.models.gnn import GNN, MSGNN
    cls = MSGNN if model_type == "MSGNN" else GNN
    m = cls(**ctor)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=lr, weight_decay=0.01)
    spec = O.ModelSpec(model_type, **ctor)
    losses = []
    for _ in range(n_steps):
        opt.zero_grad(set_to_none=True)
        g = O.adapt_batch(data_or_batch, graphs) if graphs is not None else data_or_batch
        l = O.training_step(sd, spec, g, 1, only_where_water=True, velocity_scaler=7.0)
        l.backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), 1.0)
        opt.step()
        losses.append(float(l.detach()))
    return sd, losses


@pytest.mark.parametrize("use_graph", [False, True])
def test_whole_training_step_vs_oracle_and_torch_adamw(use_graph):
    from mswe_gnn_b200.models.gnn import MSGNN
    from mswe_gnn_b200.training.train import TrainStepRunner
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, **REF_CONFIG_MODELS)
    graphs = [make_tri_mesh(16, 12, 3, seed=s, rollout_steps=1) for s in (1, 2)]
    batch = Batch.from_data_list(graphs)
    model = MSGNN(**ctor).to(DEV)
    opt = FlatAdamW(model, lr=1e-3, weight_decay=0.01, max_norm=1.0)
    runner = TrainStepRunner(model, batch.to(DEV), opt, rollout_steps=1, use_cuda_graph=False)
    ref_sd, ref_losses = _oracle_steps("MSGNN", ctor, batch, graphs, 3, 1e-3)
    if use_graph:
        # the constructor's warm-up steps already moved the weights: compare a captured runner with an eager one instead
        m2 = MSGNN(**ctor).to(DEV)
        o2 = FlatAdamW(m2, lr=1e-3, weight_decay=0.01, max_norm=1.0)
        g_runner = TrainStepRunner(m2, batch.to(DEV), o2, rollout_steps=1, use_cuda_graph=True, warmup=2)
        for _ in range(3):                                       # warm-up 2 + capture 0 + 3 replays = 5 steps ...
            lg = g_runner.step()
        for _ in range(5):                                       # ... against 5 eager steps
            le = runner.step()
        assert abs(float(lg) - float(le)) <= 1e-6 * abs(float(le)) + 1e-7
        for (k, pa), pb in zip(model.named_parameters(), m2.parameters()):
            assert rel_l2(pb.detach(), pa.detach()) < 1e-5, k
        assert float(o2.state[0]) == 5.0 and g_runner.launches_per_step > 0
        return
    ours_losses = [float(runner.step()) for _ in range(3)]
    for a, b in zip(ours_losses, ref_losses):
        assert abs(a - b) <= 5e-5 * max(1.0, abs(b)), (ours_losses, ref_losses)
    assert ours_losses[-1] < ours_losses[0]                      # and it trains
    worst = max(rel_l2(p.detach().cpu(), ref_sd[k].detach()) for k, p in model.named_parameters())
    assert worst < 5e-3, worst                                   # 3 Adam steps of lr 1e-3: sign-like updates amplify 1e-5 gradient noise
