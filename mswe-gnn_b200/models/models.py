"""Host-side mirror of the reference's ``models/models.py`` (model base class + MLP factory).

Same public names, constructor arguments, parameter names / shapes and RNG consumption order as
the reference (so ``state_dict`` round-trips with reference checkpoints and a given ``seed``
produces the same initial weights), but these modules are *parameter containers*: the arithmetic
(``Linear → activation`` stacks, residual connection ``models/models.py:50-77``, dry mask
``models/models.py:79-91``) runs inside the fused CUDA kernels (``swe_node_encode_fwd``,
``swe_edge_gate_fwd``, ``swe_decode_head_fwd`` ...).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

_ACTIVATIONS = {
    "relu": lambda dev: nn.ReLU(),
    "prelu": lambda dev: nn.PReLU(device=dev),
    "leakyrelu": lambda dev: nn.LeakyReLU(0.1),
    "elu": lambda dev: nn.ELU(),
    "swish": lambda dev: nn.SiLU(),
    "sigmoid": lambda dev: nn.Sigmoid(),
    "tanh": lambda dev: nn.Tanh(),
}


def activation_functions(activation_name, device="cpu"):
    """Activation module by name (reference ``models/models.py:149-169``)."""
    if activation_name is None:
        return None
    try:
        return _ACTIVATIONS[activation_name](device)
    except KeyError:
        raise AttributeError('Please choose one of the following options:\n'
                             '"relu", "prelu", "leakyrelu", "elu", "swish", "sigmoid", "tanh"') from None


def activation_name_of(module: Optional[nn.Module]) -> Optional[str]:
    """Inverse of :func:`activation_functions` (used when packing an MLP for the kernels)."""
    if module is None:
        return None
    for name, cls in (("relu", nn.ReLU), ("prelu", nn.PReLU), ("leakyrelu", nn.LeakyReLU), ("elu", nn.ELU),
                      ("swish", nn.SiLU), ("sigmoid", nn.Sigmoid), ("tanh", nn.Tanh)):
        if isinstance(module, cls):
            return name
    raise TypeError(f"unsupported activation module {module!r}")


def make_mlp(input_size, output_size, hidden_size=32, n_layers=2, bias=False, activation="relu", dropout=0,
             layer_norm=False, device="cpu"):
    """``nn.Sequential`` of ``n_layers`` × (Linear, activation) — the activation follows EVERY
    layer, the last one included (reference ``models/models.py:121-146``).  LayerNorm / Dropout
    variants are not implemented by the kernels and are refused here rather than silently
    computed differently."""
    if layer_norm:
        raise NotImplementedError("layer_norm=True is not supported by the B200 kernels")
    if dropout:
        raise NotImplementedError("dropout>0 is not supported by the B200 kernels")
    if n_layers < 1:
        raise ValueError("n_layers must be >= 1")
    widths = [input_size] + [hidden_size] * (n_layers - 1) + [output_size]
    mods: List[nn.Module] = []
    for fan_in, fan_out in zip(widths[:-1], widths[1:]):
        mods.append(nn.Linear(fan_in, fan_out, bias=bias, device=device))
        act = activation_functions(activation, device=device)
        if act is not None:
            mods.append(act)
    return nn.Sequential(*mods)


def init_true_residuals_weights(previous_t: int, base=2, repeat=1, device="cpu"):
    """w_t ∝ base**t, normalised, newest step heaviest (reference ``models/models.py:93-100``)."""
    w = torch.tensor([float(base ** e) for e in range(previous_t)], dtype=torch.float32, device=device)
    w = w / w.sum()
    return nn.Parameter(w.repeat(repeat).reshape(repeat, -1).T.contiguous())


class BaseFloodModel(nn.Module):
    """Seeded base class holding the temporal residual weights (reference
    ``models/models.py:7-48``).  ``learned_residuals``: True (one weight per past step shared by
    both variables), 'all' (per variable), False (add the last step un-weighted), None (nothing).
    """

    def __init__(self, previous_t=1, learned_residuals=None, seed=42, residuals_base=2, residual_init="exp",
                 with_WL=False, device="cpu"):
        super().__init__()
        torch.manual_seed(seed)
        assert residual_init in ("exp", "random"), "Argument 'residual_init' can only be either 'exp' or 'random'"
        self.previous_t = previous_t
        self.with_WL = with_WL
        self.learned_residuals = learned_residuals
        self.device = device
        self.residuals_base = residuals_base
        self.residual_init = residual_init
        self.NUM_WATER_VARS = 2
        self.out_dim = self.NUM_WATER_VARS
        repeat = {True: 1, "all": self.out_dim}.get(learned_residuals) if learned_residuals in (True, "all") else None
        if repeat is not None:
            if residual_init == "exp":
                self.residual_weights = init_true_residuals_weights(previous_t, residuals_base, repeat=repeat,
                                                                    device=device)
            else:
                self.residual_weights = nn.Parameter(torch.Tensor(previous_t, repeat).to(device))
                nn.init.xavier_normal_(self.residual_weights)

    # residual mode code understood by swe_decode_head_fwd
    def _residual_mode(self) -> int:
        lr = self.learned_residuals
        if lr is True:
            return 1
        if lr == "all":
            return 2
        if lr is False:
            return 3
        return 0

    def _apply(self, fn, *args, **kwargs):
        # keep `self.device` in step with .to()/.cuda() so callers that read model.device
        # (reference training/train.py) see where the parameters live
        out = super()._apply(fn, *args, **kwargs)
        for p in self.parameters():
            self.device = p.device
            break
        return out
