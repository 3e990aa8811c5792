"""Rollout helpers of the reference's ``utils/dataset.py`` that sit on the hot path.

Inside :func:`mswe_gnn_b200.training.train.rollout_test` these three operations are fused into
the kernels (``swe_apply_bc``; the window shift is the ``x_next`` output of
``swe_decode_head_fwd``).  The functions below keep the reference's call signatures for code that
drives a model step by step (``training/train.py:88-92``); they are a few tensor-indexing calls
on whatever device the tensors live on and contain no model arithmetic.
"""
from __future__ import annotations

import torch

NUM_WATER_VARS = 2


def check_type_BC(type_BC, num_water_vars=NUM_WATER_VARS):
    """Reference ``utils/dataset.py:499-506``."""
    if type_BC in (1, 2):
        assert type_BC <= num_water_vars, "The boundary conditions are not compatible with the data format you are using."
    elif type_BC == 3:
        raise ValueError("Vector boundary conditions are not yet implemented. "
                         "Please desist from convincing me to implement them.")
    else:
        raise ValueError(f"BC_type={type_BC} is not a valid input. Please select either:\n"
                         "1: Inflow water depth\n2: Inflow discharge")


def apply_boundary_condition(x_d, BC, node_BC, type_BC=2):
    """Write the inflow boundary condition into the dynamic columns of the ghost nodes
    (reference ``utils/dataset.py:486-497``): 1 = water depth h, 2 = discharge |q|."""
    check_type_BC(int(type_BC))
    x_d[node_BC, (int(type_BC) - 1)::NUM_WATER_VARS] = BC
    return x_d


def use_prediction(x, pred, previous_t):
    """Next-step input: static columns, window shifted left by one (h,q) pair, prediction
    appended (reference ``utils/dataset.py:508-529``)."""
    assert pred.shape[-1] == NUM_WATER_VARS, \
        "The number of predictions is not consistent with the number of future time steps"
    n_dyn = previous_t * NUM_WATER_VARS
    n_static = x.shape[1] - n_dyn
    pieces = [x[:, :n_static]] + ([x[:, n_static + NUM_WATER_VARS:]] if previous_t > 1 else []) + [pred]
    out = torch.cat(pieces, 1)
    assert out.shape == x.shape, f"The shape of the input has changed from {x.shape} to {out.shape}"
    return out


def create_scale_mask(num_nodes, num_scales, node_ptr, data_type=None, device="cpu"):
    """int32 scale id per node (reference ``utils/dataset.py:615-638``); ``node_ptr`` is ``[S+1]``
    for one graph or ``[G, S+1]`` for an adapted batch."""
    mask = torch.zeros(num_nodes, dtype=torch.int, device=device)
    rows = node_ptr.reshape(-1, node_ptr.shape[-1]).tolist()
    for s in range(num_scales):
        for r in rows:
            mask[r[s]:r[s + 1]] = s
    return mask
