"""Oracle: BYOL cosine loss and EMA target update (TEST INFRASTRUCTURE ONLY).

* ``byol_loss``   restates ref:src/models/byol.py:104-129
* ``ema_update``  restates ref:src/models/byol.py:62-73
* ``projection_head`` / ``prediction_head`` build the stock-torch heads of
  ref:src/models/multi_layer_heads.py:4-49 (used by the whole-step oracle)
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def byol_loss(online_pred: torch.Tensor, target_proj: torch.Tensor) -> torch.Tensor:
    """ref:src/models/byol.py:104-129 (the NaN checks there only log, they do not change the value)."""
    online_pred = online_pred + 1e-10                                # :113
    target_proj = target_proj + 1e-10                                # :114
    online_pred = F.normalize(online_pred, dim=1, eps=1e-10)         # :117
    target_proj = F.normalize(target_proj, dim=1, eps=1e-10)         # :118
    similarity = torch.sum(online_pred * target_proj, dim=1)         # :125
    similarity = torch.clamp(similarity, min=-1.0, max=1.0)          # :126
    return 2 - 2 * similarity.mean()                                 # :127


def ema_update(online_params, target_params, ema_decay: float):
    """ref:src/models/byol.py:62-73: t = decay*t + (1-decay)*o, per tensor, returning NEW tensors
    (the reference re-binds ``target_param.data``).  ``(1 - ema_decay)`` is evaluated in Python
    double and applied as an fp32 scalar by torch, exactly as in the reference expression."""
    out = []
    with torch.no_grad():
        for o, t in zip(online_params, target_params):
            out.append(ema_decay * t + (1 - ema_decay) * o)          # :67-68 / :72-73
    return out


def projection_head(input_dim: int, hidden_dim: int, output_dim: int) -> nn.Sequential:
    """ref:src/models/multi_layer_heads.py:15-21"""
    return nn.Sequential(
        nn.Linear(input_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
        nn.Linear(hidden_dim, output_dim), nn.BatchNorm1d(output_dim),
    )


def prediction_head(input_dim: int, hidden_dim: int, output_dim: int) -> nn.Sequential:
    """ref:src/models/multi_layer_heads.py:38-46"""
    return nn.Sequential(
        nn.Linear(input_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
        nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
        nn.Linear(hidden_dim, output_dim),
    )
