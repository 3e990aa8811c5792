"""TEST INFRASTRUCTURE — not product code.  CPU restatement (plain torch) of the dataset-side helpers next to the hot
path: temporal windows (`/root/reference/utils/dataset.py:340-471`) and rollout metrics
(`/root/reference/utils/miscellaneous.py:123-199`).  Pinned bit-exactly against the unmodified reference imported under
`oracle/ref_stubs.py` by `tests/test_dataset_oracle_pinned.py` (runs where /root/reference exists); the GPU tests compare
`swe_temporal_window` / `swe_rollout_metrics` with these functions."""
from __future__ import annotations

import torch


def add_dry_bed(v: torch.Tensor, previous_t: int) -> torch.Tensor:
    """dataset.py:376-386."""
    if v.dim() == 1:
        return torch.cat((torch.zeros(previous_t - 1), v))
    return torch.cat((torch.zeros(v.shape[0], previous_t - 1), v), 1)


def temporal_samples_size(maximum_time: int, time_start: int = 0, time_stop: int = -1, rollout_steps: int = 1) -> int:
    """dataset.py:388-407."""
    if time_stop != maximum_time:
        time_stop = time_stop % maximum_time - time_start + 1
    return time_stop - rollout_steps if rollout_steps > 0 else -rollout_steps


def temporal_sample(x_static, WD, V, BC, init_time: int, previous_t: int, rollout_steps: int):
    """One element of `to_temporal`'s list (dataset.py:409-463): (x, y, BC window)."""
    WDp, Vp = add_dry_bed(WD, previous_t), add_dry_bed(V, previous_t)
    BCp = torch.cat((add_dry_bed(BC, previous_t), BC[:, -1:]), 1)
    prev = torch.cat([torch.cat((WDp[:, s:s + 1], Vp[:, s:s + 1]), 1) for s in range(init_time, init_time + previous_t)], -1)
    nxt = torch.stack([torch.cat((WDp[:, init_time + previous_t + r:init_time + previous_t + r + 1],
                                  Vp[:, init_time + previous_t + r:init_time + previous_t + r + 1]), 1)
                       for r in range(rollout_steps)], -1)
    bc = torch.stack([BCp[:, init_time + r:init_time + r + previous_t] for r in range(rollout_steps + 1)], -1)
    return torch.cat((x_static, prev), 1), nxt, bc


def confusion(pred, real, water_threshold=0.0):
    """miscellaneous.py:123-151 for one simulation ([N, 2, T]): TP, TN, FP, FN per time step."""
    pf, rf = pred[:, 0, :] > water_threshold, real[:, 0, :] > water_threshold
    return (pf & rf).sum(0), (~pf & ~rf).sum(0), (pf & ~rf).sum(0), (~pf & rf).sum(0)


def get_CSI(pred, real, water_threshold=0.0):
    TP, TN, FP, FN = confusion(pred, real, water_threshold)
    return TP / (TP + FN + FP)


def get_F1(pred, real, water_threshold=0.0):
    TP, TN, FP, FN = confusion(pred, real, water_threshold)
    return TP / (TP + 0.5 * (FN + FP))


def get_rollout_loss(pred, real, type_loss="RMSE", only_where_water=False):
    """miscellaneous.py:177-199 for one simulation: per-variable loss, [2]."""
    d = pred - real
    if only_where_water:
        w = (d != 0).any(1)                                                 # [N, T]
        m = torch.stack([d[:, v, :][w] for v in range(d.shape[1])])         # [2, selected]
        return torch.sqrt((m ** 2).mean(-1)) if type_loss == "RMSE" else m.abs().mean(-1)
    e = torch.sqrt((d ** 2).mean(0)) if type_loss == "RMSE" else d.abs().mean(0)   # [2, T]
    return e.mean(-1)
