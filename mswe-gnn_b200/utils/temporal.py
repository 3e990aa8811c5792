"""Temporal samples and rollout metrics on the device.

``TemporalWindows`` holds one simulation (static node columns, water depth and discharge series, boundary series) on the
GPU and cuts the sample for any start time out of it with ONE kernel (``swe_temporal_window``) — the reference builds the
whole list of ``Data`` objects up front on the host (``/root/reference/utils/dataset.py:410-471``, ``to_temporal``).
``rollout_metrics`` evaluates, in one pass over a ``[N, 2, T]`` rollout, the confusion matrices behind CSI / F1
(``utils/miscellaneous.py:123-175``) and the error sums behind ``get_rollout_loss`` (``miscellaneous.py:177-199``); nothing
is read back to the host.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from .. import lib
from .data import Data


def get_temporal_samples_size(maximum_time: int, time_start: int = 0, time_stop: int = -1, rollout_steps: int = 1) -> int:
    """Reference ``utils/dataset.py:388-407`` (same assertions)."""
    assert maximum_time > 0, 'The temporal size of the dataset is zero'
    assert time_stop <= maximum_time, 'time_stop cannot be higher than the temporal size of the dataset'
    if time_stop != maximum_time:
        time_stop = time_stop % maximum_time - time_start + 1
    assert time_start <= time_stop, 'time_start cannot be higher than the last selected time'
    assert rollout_steps <= time_stop, 'Number of rollout_steps is too high'
    size = time_stop - rollout_steps if rollout_steps > 0 else -rollout_steps
    assert size >= 0, f'Something went wrong here: the temporal sample size is {size}'
    return size


class TemporalWindows:
    """``to_temporal(data, previous_t, time_start, time_stop, rollout_steps)`` without the list: ``len()`` samples,
    ``sample(i)`` / ``[i]`` cut on the device into reusable buffers (``fresh=True`` allocates new ones)."""

    def __init__(self, data, previous_t: int = 2, time_start: int = 0, time_stop: int = -1, rollout_steps: int = 1):
        if not data.WD.is_cuda:
            raise RuntimeError("TemporalWindows works on CUDA tensors only (no CPU fallback)")
        self.data, self.previous_t, self.time_start = data, int(previous_t), int(time_start)
        T = int(data.WD.shape[1])
        self.n_samples = get_temporal_samples_size(T, time_start, time_stop, rollout_steps)
        self.rollout_steps = (rollout_steps % (time_stop % T - time_start + 1)) if rollout_steps < 0 else int(rollout_steps)
        f32 = lambda t: t.to(torch.float32).contiguous()
        self.xs, self.WD, self.V, self.BC = f32(data.x), f32(data.WD), f32(data.V), f32(data.BC)
        if self.BC.dim() != 2:
            raise ValueError("data.BC must be [n_BC, T] (one boundary series per boundary node)")
        self._buf = None

    def __len__(self):
        return self.n_samples

    def _buffers(self):
        n, dev = self.WD.shape[0], self.WD.device
        return (torch.empty(n, self.xs.shape[1] + 2 * self.previous_t, device=dev),
                torch.empty(n, 2, self.rollout_steps, device=dev),
                torch.empty(self.BC.shape[0], self.previous_t, self.rollout_steps + 1, device=dev))

    def sample(self, i: int, fresh: bool = False) -> Data:
        if not 0 <= i < self.n_samples:
            raise IndexError(i)
        if fresh or self._buf is None:
            bufs = self._buffers()
            if not fresh:
                self._buf = bufs
        else:
            bufs = self._buf
        x, y, bc = bufs
        lib.temporal_window(self.xs, self.WD, self.V, self.BC, self.time_start + i, self.previous_t, self.rollout_steps, x, y, bc)
        d = self.data
        out = Data(x=x, y=y, BC=bc, time=self.time_start + i, previous_t=self.previous_t)
        for k in ("edge_index", "edge_attr", "pos", "area", "temporal_res", "edge_BC_length", "node_BC", "type_BC", "node_ptr",
                  "edge_ptr", "intra_edge_ptr", "intra_mesh_edge_index"):
            if hasattr(d, k):
                setattr(out, k, getattr(d, k))
        return out

    __getitem__ = sample


def rollout_metrics(pred: torch.Tensor, real: torch.Tensor, water_thresholds: Sequence[float] = (0.05, 0.3)) -> Dict[str, torch.Tensor]:
    """pred, real: ``[N, 2, T]`` (one simulation, as ``rollout_test`` returns).  Device tensors out:
    ``CSI`` / ``F1`` ``[n_thresholds, T]`` (``get_CSI`` / ``get_F1``), ``rmse`` / ``mae`` ``[2]`` (``get_rollout_loss``),
    ``rmse_wet`` / ``mae_wet`` ``[2]`` (``only_where_water=True``), ``confusion`` ``[n_thresholds, 4, T]`` (TP, TN, FP, FN)."""
    if pred.shape != real.shape or pred.dim() != 3 or pred.shape[1] != 2:
        raise ValueError("rollout_metrics takes [N, 2, T] tensors")
    if len(water_thresholds) > 4:
        raise ValueError("at most 4 water-depth thresholds per call")
    pred, real = pred.to(torch.float32).contiguous(), real.to(torch.float32).contiguous()
    n, _, T = pred.shape
    dev = pred.device
    K = len(water_thresholds)
    thr = torch.tensor(list(water_thresholds), dtype=torch.float32, device=dev) if K else None
    cols = lib.rollout_metrics_cols()
    out = torch.empty(T, cols, dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.load().swe_rollout_metrics_ws_bytes(T)), dtype=torch.uint8, device=dev)
    lib.rollout_metrics(pred, real, thr, out, ws)
    conf = out[:, :16].reshape(T, 4, 4).permute(1, 2, 0)[:K]               # [K, 4, T]
    TP, FP, FN = conf[:, 0], conf[:, 2], conf[:, 3]
    e = out[:, 16:]
    res = {"confusion": conf, "CSI": (TP / (TP + FN + FP)).float(), "F1": (TP / (TP + 0.5 * (FN + FP))).float()}
    res["rmse"] = torch.sqrt(e[:, 0:2] / n).mean(0).float()                # mean over time of the per-step RMSE
    res["mae"] = (e[:, 2:4] / n).mean(0).float()
    cnt = e[:, 8].sum()
    res["rmse_wet"] = torch.sqrt(e[:, 4:6].sum(0) / cnt).float()           # over all wet (node, time) entries
    res["mae_wet"] = (e[:, 6:8].sum(0) / cnt).float()
    return res
