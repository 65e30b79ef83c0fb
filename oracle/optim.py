"""Oracle: the optimizer tail of the BYOL step (TEST INFRASTRUCTURE ONLY).

Restates, with plain torch CPU ops in the reference's order, what ref:train_byol.py:67-71 executes:

* ``clip_grad_norm``  ``torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)``  (ref:train_byol.py:67;
  torch/nn/utils/clip_grad.py: per-tensor L2 norms, the L2 norm of those, ``max_norm / (total + 1e-6)`` clamped to 1,
  every gradient multiplied by it -- also when the coefficient is 1)
* ``adamw_step``      ``optimizer.step()`` of ``optim.AdamW(model.parameters(), lr, weight_decay)``
  (ref:train_byol.py:146,70; torch/optim/adam.py::_single_tensor_adam with decoupled weight decay)
* the EMA of ref:src/models/byol.py:62-73 is ``oracle.ema_update``

Pinned by ``tests/golden/optim_step.npz``: three steps of the real ``clip_grad_norm_`` + ``torch.optim.AdamW`` +
the reference's ``_update_target_network`` on a shimmed model (tests/golden/make_golden.py::gen_optim_step).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .byol import ema_update


def clip_grad_norm(grads: Sequence[Optional[torch.Tensor]], max_norm: float) -> torch.Tensor:
    """In place on ``grads``; returns the total norm (torch/nn/utils/clip_grad.py::clip_grad_norm_, norm_type 2)."""
    gs = [g for g in grads if g is not None]
    if not gs:
        return torch.tensor(0.0)
    norms = [torch.linalg.vector_norm(g, 2.0) for g in gs]
    total = torch.linalg.vector_norm(torch.stack(norms), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in gs:
        g.mul_(coef)
    return total


def adamw_step(params, grads, exp_avgs, exp_avg_sqs, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
               weight_decay: float = 1e-2) -> None:
    """In place; ``step`` is the 1-based step count.  torch/optim/adam.py::_single_tensor_adam, non-capturable path."""
    beta1, beta2 = betas
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        if g is None:
            continue
        if weight_decay != 0:
            p.mul_(1 - lr * weight_decay)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bias_correction1 = 1 - beta1 ** step
        bias_correction2 = 1 - beta2 ** step
        step_size = lr / bias_correction1
        bias_correction2_sqrt = bias_correction2 ** 0.5
        denom = (v.sqrt() / bias_correction2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def clip_adamw_ema_step(params, grads, exp_avgs, exp_avg_sqs, targets, step: int, lr: float, betas=(0.9, 0.999),
                        eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 1.0,
                        ema_decay: float = 0.996) -> torch.Tensor:
    """ref:train_byol.py:67-71 on lists of CPU tensors (in place; ``targets[i]`` may be None).  Returns the norm."""
    with torch.no_grad():
        total = clip_grad_norm(grads, max_norm) if max_norm > 0 else torch.tensor(0.0)
        adamw_step(params, grads, exp_avgs, exp_avg_sqs, step, lr, betas, eps, weight_decay)
        idx = [i for i, t in enumerate(targets) if t is not None]
        new = ema_update([params[i] for i in idx], [targets[i] for i in idx], ema_decay)
        for i, t in zip(idx, new):
            targets[i].copy_(t)
    return total
