"""Loss of the training step — host-side mirror of the reference's ``training/loss.py:76-118``
(``loss_function`` with ``conservation = 0``; the mass-conservation term needs the dataset's mesh
areas and is out of scope).  A handful of torch ops on the ``[N, 2]`` predictions:
autograd turns it into ``d loss / d pred`` and hands that to the backward kernels."""
from __future__ import annotations

import torch

NUM_WATER_VARS = 2
_SCALER_CACHE = {}


def get_mean_error(diff_rollout, type_loss, nodes_dim=0):
    if type_loss == 'RMSE':
        return torch.sqrt((diff_rollout ** 2).mean(nodes_dim))
    if type_loss == 'MAE':
        return diff_rollout.abs().mean(nodes_dim)
    raise ValueError("loss_type must be either 'RMSE' or 'MAE'")


def mask_on_water(diff, water_axis=1):
    return (diff != 0).any(water_axis)


def get_loss_variable_scaler(velocity_scaler=1):
    loss_scaler = torch.ones(NUM_WATER_VARS)
    loss_scaler[1::NUM_WATER_VARS] = velocity_scaler
    return loss_scaler


def _masked_mean_error(diff, keep, type_loss):
    """``get_mean_error(diff[keep])`` without the data-dependent shape (boolean indexing makes the host wait for the
    device): Σ over kept rows / number of kept rows.  Same value up to fp32 summation order."""
    w = keep.to(diff.dtype).unsqueeze(1)
    cnt = w.sum()
    if type_loss == 'RMSE':
        return torch.sqrt((diff * diff * w).sum(0) / cnt)
    if type_loss == 'MAE':
        return (diff.abs() * w).sum(0) / cnt
    raise ValueError("loss_type must be either 'RMSE' or 'MAE'")


def get_multiscale_loss(diff, data, only_where_water=True, type_loss='RMSE', nodes_dim=0):
    """Finest-scale rows only (reference ``loss.py:49-74``)."""
    node_ptr = data.node_ptr
    finest = getattr(data, "_finest_rows", None)           # cached by training.train._adapt_cached: no host reads
    if finest is not None and finest.device == diff.device and nodes_dim == 0:
        keep = finest & mask_on_water(diff) if only_where_water else finest
        return _masked_mean_error(diff, keep, type_loss)
    where_water = mask_on_water(diff) if only_where_water else torch.ones(diff.shape[0], dtype=torch.bool, device=diff.device)
    if node_ptr.dim() == 2:
        ptr = node_ptr.tolist()
        parts = [diff[p[0]:p[1]][where_water[p[0]:p[1]]] for p in ptr]
        return get_mean_error(torch.cat(parts), type_loss, nodes_dim)
    lo, hi = int(node_ptr[0]), int(node_ptr[1])
    return get_mean_error(diff[lo:hi][where_water[lo:hi]], type_loss, nodes_dim)


def loss_function(preds, real, data, BC=None, type_loss='RMSE', only_where_water=False, conservation=0, velocity_scaler=1):
    if conservation != 0:
        raise NotImplementedError("the mass-conservation loss term is outside the B200 hot path")
    diff = preds - real
    if 'node_ptr' in data.keys():
        loss = get_multiscale_loss(diff, data, only_where_water, type_loss, nodes_dim=0)
    elif only_where_water and getattr(data, "_finest_rows", None) is not None:
        loss = _masked_mean_error(diff, mask_on_water(diff), type_loss)
    else:
        if only_where_water:
            diff = diff[mask_on_water(diff)]
        loss = get_mean_error(diff, type_loss, nodes_dim=0)
    key = (str(diff.device), float(velocity_scaler))
    loss_scaler = _SCALER_CACHE.get(key)
    if loss_scaler is None:       # a host -> device copy from pageable memory makes the host wait for the stream: once only
        loss_scaler = get_loss_variable_scaler(velocity_scaler=velocity_scaler).to(diff.device)
        _SCALER_CACHE[key] = loss_scaler
    return torch.dot(loss, loss_scaler) / loss_scaler.sum()
