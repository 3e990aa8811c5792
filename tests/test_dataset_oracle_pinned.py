"""CPU, only where /root/reference exists: oracle/dataset_oracle.py against the UNMODIFIED reference's `to_temporal` and
metric helpers imported under the stubs (bit-exact)."""
import importlib

import pytest
import torch

from oracle import dataset_oracle as DO
from oracle import ref_stubs

pytestmark = pytest.mark.skipif(not ref_stubs.reference_available(), reason="/root/reference not present")


def _sim(n=37, T=11, n_bc=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    WD = torch.rand(n, T, generator=g) * (torch.rand(n, 1, generator=g) < 0.6)
    V = torch.rand(n, T, generator=g) * (WD > 0)
    return torch.rand(n, 3, generator=g), WD, V, torch.rand(n_bc, T, generator=g)


@pytest.mark.parametrize("previous_t,rollout_steps,time_start,time_stop", [(2, 1, 0, -1), (3, 4, 0, -1), (1, 2, 2, 8), (3, 1, 1, -1)])
def test_temporal_samples_bit_exact(previous_t, rollout_steps, time_start, time_stop):
    ref_stubs.install()
    ds = importlib.import_module("utils.dataset")
    xs, WD, V, BC = _sim()
    data = ref_stubs.StubData(x=xs, WD=WD, V=V, BC=BC, edge_index=torch.zeros(2, 0, dtype=torch.long), edge_attr=None, pos=None,
                              area=None, temporal_res=60, edge_BC_length=None, node_BC=torch.tensor([0, 1]), type_BC=2)
    ref = ds.to_temporal(data, previous_t=previous_t, time_start=time_start, time_stop=time_stop, rollout_steps=rollout_steps)
    assert len(ref) == DO.temporal_samples_size(WD.shape[1], time_start, time_stop, rollout_steps)
    for i, r in enumerate(ref):
        x, y, bc = DO.temporal_sample(xs, WD, V, BC, time_start + i, previous_t, rollout_steps)
        assert torch.equal(x, r.x) and torch.equal(y, r.y) and torch.equal(bc, r.BC) and r.time == time_start + i


def test_metrics_bit_exact():
    ref_stubs.install()
    import sys
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.animation", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    wb = types.ModuleType("wandb")                               # `from wandb import Config` at the top of that module
    wb.Config = dict
    saved = sys.modules.get("wandb")
    sys.modules["wandb"] = wb
    try:
        misc = importlib.import_module("utils.miscellaneous")
    except Exception as e:                                       # plotting / logging dependencies of that module
        pytest.skip(f"utils.miscellaneous not importable here: {e}")
    finally:
        if saved is not None:
            sys.modules["wandb"] = saved
        else:
            sys.modules.pop("wandb", None)
    g = torch.Generator().manual_seed(1)
    real = torch.rand(50, 2, 7, generator=g) * (torch.rand(50, 1, 7, generator=g) < 0.5)
    pred = (real + 0.1 * torch.randn(50, 2, 7, generator=g)).clamp_min(0) * (torch.rand(50, 1, 7, generator=g) < 0.8)
    for thr in (0.0, 0.05, 0.3):
        assert torch.equal(DO.get_CSI(pred, real, thr), misc.get_CSI(pred, real, thr))
        assert torch.equal(DO.get_F1(pred, real, thr), misc.get_F1(pred, real, thr))
    for tl in ("RMSE", "MAE"):
        for oww in (False, True):
            assert torch.equal(DO.get_rollout_loss(pred, real, tl, oww), misc.get_rollout_loss(pred, real, tl, oww))
