"""The dimensional-emotion fine-tune step (BASELINE.json configs[3]) with the reference's pieces
(ref:src/train/dimentional_emotions.py:306-465): ``ccc_loss``, ``compute_ccc``, the step body of
``train_one_epoch_dimensional`` and the epoch loop.

Differences underneath: the CCC loss is vectorised over the three dimensions (one reduction instead of a Python loop of
~10 ops per dimension); the optimizer tail can be ``FusedAdamWEma`` (clip + AdamW in two launches, no EMA on this
path); predictions are accumulated on the device and read back once per epoch instead of 6 ``.cpu().numpy()`` per step
(ref:...:343-348)."""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import numpy as np
import torch

from .optim import FusedAdamWEma


def ccc_loss(predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """mean over dimensions of (1 - CCC), population variances, eps 1e-10; 0 for a batch of one
    (ref:src/train/dimentional_emotions.py:427-450)."""
    if predictions.size(0) <= 1:
        return predictions.sum() * 0.0
    mp, mt = predictions.mean(0), targets.mean(0)
    dp, dt = predictions - mp, targets - mt
    var_p, var_t = (dp * dp).mean(0), (dt * dt).mean(0)
    covar = (dp * dt).mean(0)
    ccc = 2 * covar / (var_p + var_t + (mp - mt) ** 2 + 1e-10)
    return (1 - ccc).sum() / predictions.size(1)


def compute_ccc(predictions, targets) -> float:
    """ref:src/train/dimentional_emotions.py:453-465 (numpy, whole epoch)."""
    predictions, targets = np.asarray(predictions), np.asarray(targets)
    mp, mt = np.mean(predictions), np.mean(targets)
    covar = np.mean((predictions - mp) * (targets - mt))
    return float(2 * covar / (np.var(predictions) + np.var(targets) + (mp - mt) ** 2 + 1e-10))


def emotion_dim_step(model, inputs: torch.Tensor, labels: torch.Tensor, optimizer, attention_mask=None,
                     max_grad_norm: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """forward(task='dimensional') -> ccc_loss -> zero_grad -> backward -> clip_grad_norm_(1.0) -> optimizer.step
    (ref:src/train/dimentional_emotions.py:326-338).  Returns (detached loss, detached predictions), no host sync."""
    if attention_mask is None:  # :319-323
        attention_mask = torch.ones(inputs.size(0), inputs.size(-1), device=inputs.device)
    _, values = model(inputs, attention_mask=attention_mask, task="dimensional")
    loss = ccc_loss(values, labels)
    optimizer.zero_grad() if isinstance(optimizer, FusedAdamWEma) else optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if not (isinstance(optimizer, FusedAdamWEma) and optimizer.max_grad_norm > 0):
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_grad_norm)
    optimizer.step()
    return loss.detach(), values.detach()


def train_one_epoch_dimensional(model, dataloader: Iterable, optimizer, device) -> Tuple[float, Dict[str, float]]:
    """(mean loss, {'A','V','D','avg'} CCC over the epoch), ref:src/train/dimentional_emotions.py:306-358."""
    device = torch.device(device)
    model.train()
    total = torch.zeros((), device=device)
    preds, labs, n = [], [], 0
    for batch in dataloader:
        inputs = batch["input_values"].to(device, non_blocking=True)
        labels = torch.stack([batch["A"], batch["V"], batch["D"]], dim=1).to(device, non_blocking=True).float()
        mask = batch["attention_mask"].to(device, non_blocking=True) if "attention_mask" in batch else None
        loss, values = emotion_dim_step(model, inputs, labels, optimizer, mask)
        total += loss
        preds.append(values)
        labs.append(labels)
        n += 1
    if n == 0:
        return 0.0, {"A": 0.0, "V": 0.0, "D": 0.0, "avg": 0.0}
    p, l = torch.cat(preds).cpu().numpy(), torch.cat(labs).cpu().numpy()  # one read-back per epoch
    ccc = {k: compute_ccc(p[:, i], l[:, i]) for i, k in enumerate("AVD")}
    ccc["avg"] = (ccc["A"] + ccc["V"] + ccc["D"]) / 3
    return float(total.item()) / n, ccc
