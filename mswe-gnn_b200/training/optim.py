"""Optimizer step and loss of the training loop on the device.

``FlatAdamW`` keeps all parameters of a model in ONE flat fp32 buffer (every ``nn.Parameter`` becomes a view into it, the
gradients are views into a second flat buffer) and performs ``clip_grad_norm_`` + ``torch.optim.AdamW.step`` as two
kernels (``swe_clip_adamw_step``); the flat gradient is also what the data-parallel all-reduce ships.  Reference:
``/root/reference/training/train.py:147-155`` (AdamW, StepLR) and ``/root/reference/main.py:109`` (gradient_clip_val=1).

``device_loss`` is ``training/loss.py:loss_function`` (conservation = 0) as one ``torch.autograd.Function`` whose forward
computes the loss AND d loss / d pred in two kernels (``swe_loss_fwd_bwd``) — no boolean indexing, no host reads.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import lib


class FlatAdamW:
    """AdamW (+ global-norm clipping) over a flat view of ``model.parameters()``; same update rule and hyper-parameter
    names as ``torch.optim.AdamW`` (no amsgrad, no maximize)."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_norm: float = 0.0):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW runs on CUDA parameters only (no CPU fallback)")
        if any(p.dtype != torch.float32 or p.device != dev for p in params):
            raise TypeError("all parameters must be float32 tensors on one CUDA device")
        self.params = params
        # every parameter starts on a 128-byte boundary of the flat buffer (the kernels read weights with 16-byte loads);
        # the padding elements stay zero in the parameters, the gradient and both moments
        al = lambda k: (k + 31) // 32 * 32
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += al(p.numel())
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(params, self.offsets):
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)              # the parameter IS a slice of the flat buffer now
                p.grad = self.grad[off:off + k].view_as(p)              # autograd accumulates in place into the flat gradient
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)    # step count, last grad norm, last clip coefficient
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, float(eps), float(weight_decay), float(max_norm)
        self._ws = torch.empty(lib.train_step_ws_bytes(), dtype=torch.uint8, device=dev)

    def zero_grad(self):
        """Gradients stay allocated (views of the flat buffer): zeroed in place, never set to None."""
        self.grad.zero_()
        for p in self.params:                                           # someone may have replaced .grad (set_to_none)
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or \
                    p.grad.data_ptr() >= self.grad.data_ptr() + self.grad.numel() * 4:
                self._rebind_grads()
                break

    def _rebind_grads(self):
        for p, off in zip(self.params, self.offsets):
            p.grad = self.grad[off:off + p.numel()].view_as(p)

    def set_lr(self, lr: float):
        """A scheduler's new learning rate (device scalar: a captured step sees it on its next replay)."""
        self.lr.fill_(float(lr))

    def step(self):
        lib.clip_adamw_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                            self.eps, self.weight_decay, self.max_norm, self.state, self._ws)
        self.mark_updated()

    def mark_updated(self):
        """The kernel wrote the parameters through raw pointers: tell torch (and with it the packed-weight caches of
        engine.py, which are keyed on the parameters' version counters) that they changed."""
        for p in self.params:
            torch.autograd.graph.increment_version(p)

    @property
    def last_grad_norm(self) -> torch.Tensor:
        return self.state[1]


class _DeviceLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, real, real_stride, rows, only_where_water, mae, w0, w1, scale, ws):
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        dpred = torch.empty_like(pred)
        lib.loss_fwd_bwd(pred, real, real_stride, rows, pred.shape[0], only_where_water, mae, w0, w1, scale, False, loss, dpred, ws)
        ctx.save_for_backward(dpred)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None, None, None, None, None, None, None, None, None


_WS = {}


def device_loss(preds: torch.Tensor, real: torch.Tensor, rows: Optional[torch.Tensor] = None, type_loss: str = "RMSE",
                only_where_water: bool = False, velocity_scaler: float = 1.0, scale: float = 1.0) -> torch.Tensor:
    """``loss_function(preds, real, data, type_loss=..., only_where_water=..., velocity_scaler=...)`` of the reference
    (``training/loss.py:76-118``) on the device.  ``real``: ``[N, 2]`` — possibly a time slice ``y[:, :, t]`` of a
    ``[N, 2, T]`` target (read in place through its strides); ``rows``: bool ``[N]``, the finest-scale rows of a multi-scale
    graph (``loss.py:49-74``), None = all rows."""
    if type_loss not in ("RMSE", "MAE"):
        raise ValueError("loss_type must be either 'RMSE' or 'MAE'")
    if preds.dim() != 2 or preds.shape[1] != 2 or real.shape != preds.shape:
        raise ValueError("device_loss handles [N, 2] predictions (water depth, discharge)")
    if real.dtype != torch.float32 or real.stride(1) * 2 != real.stride(0):
        real = real.contiguous()
    preds = preds.contiguous()
    ws = _WS.get(preds.device)
    if ws is None:
        ws = _WS[preds.device] = torch.empty(lib.train_step_ws_bytes(), dtype=torch.uint8, device=preds.device)
    if rows is not None:
        rows = rows.view(torch.uint8) if rows.dtype == torch.bool else rows
    return _DeviceLoss.apply(preds, real, int(real.stride(0)), rows, bool(only_where_water), type_loss == "MAE", 1.0,
                             float(velocity_scaler), float(scale), ws)
