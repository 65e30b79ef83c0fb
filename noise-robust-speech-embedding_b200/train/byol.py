"""The BYOL step and its validation passes with the reference's call signatures
(ref:train_byol.py:20-79 ``train_one_epoch``, ref:evaluate_byol.py:12-123).

What is different underneath:
* batches arrive already mixed/normalised on the device (``MixedBatchLoader``); ``.to(device)`` is a no-op there and
  the reference's CPU-tensor batches still work;
* ``check_audio_tensor`` (ref:src/utils/debugging_utils.py:4-30, >= 4 host syncs per tensor, called 4x per step) is ONE
  fused kernel launch per group of tensors (``ops.check_tensors``: one status record per tensor) and ONE host read per
  checked step -- and only every ``check_interval`` steps;
* ``byol_loss`` and the EMA update are single kernel launches (ops.byol_loss, ops.EmaPlan);
* the loss is accumulated on the device; ``.item()`` is called once per epoch, not once per step;
* the per-item ``.item()`` loop of ``evaluate_embedding_similarity`` is a per-SNR masked mean on the device.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from .. import ops
from ..models.byol import BYOLSpeechModel, byol_loss, log_loss_flags
from ..utils.logging_utils import logger
from .optim import FusedAdamWEma


_CHECK_MESSAGES = ((ops.CHECK_NAN, "NaN values"), (ops.CHECK_INF, "Inf values"), (ops.CHECK_SMALL, "very small values"),
                   (ops.CHECK_LARGE, "very large values"))


def _report_checks(infos, names, config) -> list:
    """Log lines and verdicts of ref:src/utils/debugging_utils.py:4-30 from decoded check records."""
    verdicts = []
    debug = config.get("logging", {}).get("level") == "DEBUG"
    for info, name in zip(infos, names):
        bad = next((what for bit, what in _CHECK_MESSAGES if info["flags"] & bit), None)
        if bad is not None:
            if bad == "very large values":
                logger.warning("WARNING: %s contains very large values! Max abs value: %s", name, info["abs_max"])
            else:
                logger.warning("WARNING: %s contains %s!", name, bad)
            verdicts.append(False)
            continue
        if debug:
            logger.debug("Stats for %s: mean=%.4f, std=%.4f, min=%.4f, max=%.4f", name, info["mean"], info["std"],
                         info["min"], info["max"])
        verdicts.append(True)
    return verdicts


def check_audio_tensors(named, config, max_threshold: float = 1e6, min_threshold: float = 1e-6) -> list:
    """``check_audio_tensor`` for up to 8 (tensor, name) pairs at once: ONE kernel launch reads every tensor once and
    writes one status record per tensor, ONE device-to-host copy brings the records back.  Same verdicts and log lines as
    the reference (ref:src/utils/debugging_utils.py:4-30), which spends >= 4 passes and >= 4 host synchronisations per
    tensor, four tensors per step (ref:train_byol.py:52-59)."""
    records = ops.check_tensors([t for t, _ in named], max_threshold, min_threshold)
    return _report_checks(ops.decode_tensor_checks(records.cpu()), [n for _, n in named], config)


def check_audio_tensor(tensor: torch.Tensor, name: str, config, max_threshold: float = 1e6,
                       min_threshold: float = 1e-6) -> bool:
    """Drop-in for ref:src/utils/debugging_utils.py:4-30: one fused launch, one host read."""
    return check_audio_tensors([(tensor, name)], config, max_threshold, min_threshold)[0]


def byol_step(model: BYOLSpeechModel, clean: torch.Tensor, noisy: torch.Tensor, optimizer, scheduler=None,
              max_grad_norm: float = 1.0, check_config=None) -> torch.Tensor:
    """forward -> byol_loss -> zero_grad -> backward -> clip_grad_norm_(1.0) -> optimizer.step -> EMA -> scheduler.step
    (ref:train_byol.py:56-74).  Returns the detached loss (device tensor; no host sync).

    With a ``FusedAdamWEma`` optimizer that has the clip and the EMA attached (``FusedAdamWEma.for_byol``) the three
    tail stages are its two kernel launches; with any other optimizer they run as in the reference.

    ``check_config`` (the config dict) switches the reference's diagnostics on for this step (ref:train_byol.py:52-59,
    ref:src/models/byol.py:109-122): the four ``check_audio_tensor`` verdicts and the two NaN messages of ``byol_loss``,
    from two fused check launches + the loss kernel's flag word, read back together in ONE host synchronisation."""
    rec_in = ops.check_tensors([clean, noisy]) if check_config is not None else None
    online_pred, target_proj = model(clean, noisy)
    if check_config is not None:
        rec_out = ops.check_tensors([online_pred, target_proj])
        loss, flags = ops.byol_loss_with_flags(online_pred, target_proj)
        packed = torch.cat([rec_in.flatten(), rec_out.flatten(), flags.reshape(1).view(torch.uint8)]).cpu()  # the one sync
        _report_checks(ops.decode_tensor_checks(packed[:256].view(4, 64)),
                       ["clean_input_values", "noisy_input_values", "online_pred", "target_proj"], check_config)
        log_loss_flags(int(packed[256:260].view(torch.int32).item()))
    else:
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
    diagnostics -- its four ``check_audio_tensor`` calls and the NaN messages of ``byol_loss`` -- every that many steps
    at the price of ONE host synchronisation in such a step (the reference runs them every step and pays >= 20);
    0 = never: no host sync inside the epoch."""
    device = torch.device(device)
    model.train()
    total = torch.zeros((), device=device)
    n = 0
    for step, batch in enumerate(dataloader):
        clean = batch["clean_input_values"].to(device, non_blocking=True)
        noisy = batch["noisy_input_values"].to(device, non_blocking=True)
        checked = bool(check_interval) and step % check_interval == 0
        total += byol_step(model, clean, noisy, optimizer, scheduler, check_config=config if checked else None)
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
        sim = ops.cosine_rows_plain(ce.float(), ne.float()).double()  # F.normalize(dim=1) + row dot (no clamp), one launch
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
