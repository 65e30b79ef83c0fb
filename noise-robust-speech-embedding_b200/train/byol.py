"""The BYOL step and its validation passes with the reference's call signatures
(ref:train_byol.py:20-79 ``train_one_epoch``, ref:evaluate_byol.py:12-123).

What is different underneath:
* batches arrive already mixed/normalised on the device (``MixedBatchLoader``); ``.to(device)`` is a no-op there and
  the reference's CPU-tensor batches still work;
* ``check_audio_tensor`` (ref:src/utils/debugging_utils.py:4-30, >= 4 host syncs per tensor, called 4x per step) is
  evaluated as ONE fused flag vector per step and read back once -- and only every ``check_interval`` steps;
* ``byol_loss`` and the EMA update are single kernel launches (ops.byol_loss, ops.EmaPlan);
* the loss is accumulated on the device; ``.item()`` is called once per epoch, not once per step;
* the per-item ``.item()`` loop of ``evaluate_embedding_similarity`` is a per-SNR masked mean on the device.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from .. import ops
from ..models.byol import BYOLSpeechModel, byol_loss
from ..utils.logging_utils import logger
from .optim import FusedAdamWEma


def _flags(t: torch.Tensor, max_threshold: float, min_threshold: float) -> torch.Tensor:
    a = t.detach().abs()
    return torch.stack([torch.isnan(t).any(), torch.isinf(t).any(), a.sum() < min_threshold, a.max() > max_threshold])


def check_audio_tensor(tensor: torch.Tensor, name: str, config, max_threshold: float = 1e6,
                       min_threshold: float = 1e-6) -> bool:
    """Same verdicts and log lines as ref:src/utils/debugging_utils.py:4-30, with one host read instead of >= 4."""
    f = _flags(tensor, max_threshold, min_threshold).tolist()
    for bad, what in zip(f, ("NaN values", "Inf values", "very small values", "very large values")):
        if bad:
            logger.warning("WARNING: %s contains %s!", name, what)
            return False
    if config.get("logging", {}).get("level") == "DEBUG":
        t = tensor.detach().float()
        stats = torch.stack([t.mean(), t.std(), t.min(), t.max()]).tolist()
        logger.debug("Stats for %s: mean=%.4f, std=%.4f, min=%.4f, max=%.4f", name, *stats)
    return True


def byol_step(model: BYOLSpeechModel, clean: torch.Tensor, noisy: torch.Tensor, optimizer, scheduler=None,
              max_grad_norm: float = 1.0) -> torch.Tensor:
    """forward -> byol_loss -> zero_grad -> backward -> clip_grad_norm_(1.0) -> optimizer.step -> EMA -> scheduler.step
    (ref:train_byol.py:56-74).  Returns the detached loss (device tensor; no host sync).

    With a ``FusedAdamWEma`` optimizer that has the clip and the EMA attached (``FusedAdamWEma.for_byol``) the three
    tail stages are its two kernel launches; with any other optimizer they run as in the reference."""
    online_pred, target_proj = model(clean, noisy)
    loss = byol_loss(online_pred, target_proj)
    fused = isinstance(optimizer, FusedAdamWEma)
    optimizer.zero_grad() if fused else optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if not (fused and optimizer.max_grad_norm > 0):
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_grad_norm)
    optimizer.step()
    if not (fused and optimizer.has_ema):
        inner = model.module if hasattr(model, "module") else model  # DistributedDataParallel wrapper
        inner._update_target_network()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()


def train_one_epoch(model, dataloader: Iterable, optimizer, scheduler, device, config, check_interval: int = 0) -> float:
    """Average training loss over the epoch (ref:train_byol.py:20-79).  ``check_interval`` > 0 runs the reference's
    four ``check_audio_tensor`` calls every that many steps (0 = never: no host sync inside the epoch)."""
    device = torch.device(device)
    model.train()
    total = torch.zeros((), device=device)
    n = 0
    for step, batch in enumerate(dataloader):
        clean = batch["clean_input_values"].to(device, non_blocking=True)
        noisy = batch["noisy_input_values"].to(device, non_blocking=True)
        if check_interval and step % check_interval == 0:
            check_audio_tensor(clean, "clean_input_values", config)
            check_audio_tensor(noisy, "noisy_input_values", config)
        total += byol_step(model, clean, noisy, optimizer, scheduler)
        n += 1
    return float(total.item()) / max(n, 1)


@torch.no_grad()
def evaluate_embedding_similarity(model, dataloader: Iterable, device, config) -> Dict[int, float]:
    """{snr: mean cosine similarity between clean and noisy embeddings} (ref:evaluate_byol.py:12-66)."""
    device = torch.device(device)
    inner = model.module if hasattr(model, "module") else model
    inner.eval()
    encoder = inner.get_encoder()
    snr_range = list(config["data"]["snr_range"])
    snr_t = torch.tensor(snr_range, device=device)
    sums = torch.zeros(len(snr_range), device=device, dtype=torch.float64)
    counts = torch.zeros(len(snr_range), device=device, dtype=torch.float64)
    for batch in dataloader:
        clean = batch["clean_input_values"].to(device, non_blocking=True)
        noisy = batch["noisy_input_values"].to(device, non_blocking=True)
        snr = torch.as_tensor(batch["snr"]).to(device)
        ce, ne = inner._pool(encoder(clean)), inner._pool(encoder(noisy))
        sim = ops.cosine_rows(ce.float(), ne.float()).double()  # F.normalize(dim=1) + row dot, one launch
        onehot = (snr[:, None] == snr_t[None, :]).double()
        sums += (onehot * sim[:, None]).sum(0)
        counts += onehot.sum(0)
    avg = torch.where(counts > 0, sums / counts.clamp_min(1), torch.zeros_like(sums)).tolist()
    return {snr: float(v) for snr, v in zip(snr_range, avg)}


@torch.no_grad()
def validate_model(model, val_loader: Iterable, device, config) -> Tuple[float, dict]:
    """(val_loss, {'val_loss', 'val_avg_similarity', 'val_similarities'}) as in ref:evaluate_byol.py:69-123."""
    device = torch.device(device)
    inner = model.module if hasattr(model, "module") else model
    inner.eval()
    similarities = evaluate_embedding_similarity(model, val_loader, device, config)
    total = torch.zeros((), device=device)
    n = 0
    for batch in val_loader:
        clean = batch["clean_input_values"].to(device, non_blocking=True)
        noisy = batch["noisy_input_values"].to(device, non_blocking=True)
        online_pred, target_proj = inner(clean, noisy)
        total += byol_loss(online_pred, target_proj)
        n += 1
    val_loss = float(total.item()) / n if n else float("inf")
    avg_sim = sum(similarities.values()) / len(similarities) if similarities else 0.0
    return val_loss, {"val_loss": val_loss, "val_avg_similarity": avg_sim, "val_similarities": similarities}


class EarlyStopping:
    """ref:train_byol.py:82-116."""

    def __init__(self, patience: int = 5, min_delta: float = 0.0, mode: str = "min"):
        self.patience, self.min_delta, self.mode = patience, min_delta, mode
        self.counter = 0
        self.best_score: Optional[float] = None
        self.early_stop = False

    def __call__(self, score: float) -> bool:
        if self.best_score is None:
            self.best_score = score
            return False
        gain = self.best_score - score if self.mode == "min" else score - self.best_score
        if gain > self.min_delta:
            self.best_score, self.counter = score, 0
        else:
            self.counter += 1
            self.early_stop = self.counter >= self.patience
        return self.early_stop
