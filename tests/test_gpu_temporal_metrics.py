"""GPU: temporal samples (swe_temporal_window) bit-exact against the oracle restatement of the reference's `to_temporal`;
rollout metrics (swe_rollout_metrics) against the oracle's CSI / F1 / rollout loss."""
import pytest
import torch

from mswe_gnn_b200.utils.data import Data
from mswe_gnn_b200.utils.temporal import TemporalWindows, rollout_metrics
from oracle import dataset_oracle as DO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("previous_t,rollout_steps,time_start,time_stop,n", [(2, 1, 0, -1, 37), (3, 4, 0, -1, 1000), (1, 2, 2, 8, 5),
                                                                             (3, 1, 1, -1, 70001)])
def test_temporal_windows_bit_exact(previous_t, rollout_steps, time_start, time_stop, n):
    g = torch.Generator().manual_seed(n)
    T = 11
    WD = torch.rand(n, T, generator=g) * (torch.rand(n, 1, generator=g) < 0.6)
    V = torch.rand(n, T, generator=g) * (WD > 0)
    xs, BC = torch.rand(n, 3, generator=g), torch.rand(2, T, generator=g)
    d = Data(x=xs.to(DEV), WD=WD.to(DEV), V=V.to(DEV), BC=BC.to(DEV), node_BC=torch.tensor([0, 1], device=DEV), type_BC=2,
             edge_index=torch.zeros(2, 0, dtype=torch.long, device=DEV))
    tw = TemporalWindows(d, previous_t, time_start, time_stop, rollout_steps)
    assert len(tw) == DO.temporal_samples_size(T, time_start, time_stop, rollout_steps)
    for i in range(len(tw)):
        s = tw[i]
        x, y, bc = DO.temporal_sample(xs, WD, V, BC, time_start + i, previous_t, rollout_steps)
        assert torch.equal(s.x.cpu(), x) and torch.equal(s.y.cpu(), y) and torch.equal(s.BC.cpu(), bc)
        assert s.time == time_start + i and s.node_BC is d.node_BC


@pytest.mark.parametrize("n,T", [(50, 7), (100000, 48), (3, 1)])
def test_rollout_metrics_vs_oracle(n, T):
    g = torch.Generator().manual_seed(T)
    real = torch.rand(n, 2, T, generator=g) * (torch.rand(n, 1, T, generator=g) < 0.5)
    pred = (real + 0.1 * torch.randn(n, 2, T, generator=g)).clamp_min(0) * (torch.rand(n, 1, T, generator=g) < 0.8)
    thr = (0.0, 0.05, 0.3)
    m = rollout_metrics(pred.to(DEV), real.to(DEV), thr)
    for k, t in enumerate(thr):
        TP, TN, FP, FN = DO.confusion(pred, real, t)
        assert torch.equal(m["confusion"][k].cpu().long(), torch.stack([TP, TN, FP, FN]))       # integer work: exact
        a, b = m["CSI"][k].cpu(), DO.get_CSI(pred, real, t)
        assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.allclose(a.nan_to_num(), b.nan_to_num(), rtol=1e-6)
        a, b = m["F1"][k].cpu(), DO.get_F1(pred, real, t)
        assert torch.allclose(a.nan_to_num(), b.nan_to_num(), rtol=1e-6)
    assert torch.allclose(m["rmse"].cpu(), DO.get_rollout_loss(pred, real, "RMSE", False), rtol=2e-6)
    assert torch.allclose(m["mae"].cpu(), DO.get_rollout_loss(pred, real, "MAE", False), rtol=2e-6)
    assert torch.allclose(m["rmse_wet"].cpu(), DO.get_rollout_loss(pred, real, "RMSE", True), rtol=2e-6)
    assert torch.allclose(m["mae_wet"].cpu(), DO.get_rollout_loss(pred, real, "MAE", True), rtol=2e-6)
