"""``BYOLSpeechModel`` / ``byol_loss`` with the reference's surface (ref:src/models/byol.py), on the B200 kernels.

Differences, all deliberate:
* the encoder output is mean-pooled over time before the projector (``model.pooling``, default "mean").  At the
  reference's HEAD the un-pooled ``[B,T,H]`` tensor is fed to ``BatchNorm1d`` and raises; the
  logged runs used a pooled ``[B,H]`` embedding (ref:dev.ipynb:227).
* ``_update_target_network`` is ONE multi-tensor kernel launch over a chunk table built once, updating the target
  parameters in place with the reference's rounding (ref:src/models/byol.py:62-73 issues 1488 launches).
* ``byol_loss`` is one fused kernel forward and one backward, with no host synchronisation (the reference's four
  ``isnan().any()`` checks only log; use ``check_finite=True`` to get the same log lines at the cost of a sync).
State-dict keys and shapes are those of the reference (sub-module names below are the checkpoint contract).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from .. import ops
from ..utils.logging_utils import logger
from .encoder import WavLMEncoder
from .multi_layer_heads import PredictionHead, ProjectionHead


class BYOLSpeechModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        mcfg = config["model"]
        model_name = mcfg["name"]
        projection_dim = mcfg["projection_dim"]
        prediction_hidden_dim = mcfg["prediction_dim"]
        self.ema_decay = mcfg["ema_decay"]
        self.pooling = mcfg.get("pooling", "mean")
        frontend = mcfg.get("frontend", "b200")

        self.online_encoder = WavLMEncoder(model_name, frontend=frontend)
        self.online_projector = ProjectionHead(self.online_encoder.output_dim, projection_dim, projection_dim)
        self.online_predictor = PredictionHead(projection_dim, prediction_hidden_dim, projection_dim)
        self.target_encoder = WavLMEncoder(model_name, frontend=frontend)
        self.target_projector = ProjectionHead(self.target_encoder.output_dim, projection_dim, projection_dim)

        self._copy_weights(self.online_encoder, self.target_encoder)
        self._copy_weights(self.online_projector, self.target_projector)
        for p in self.target_encoder.parameters():
            p.requires_grad = False
        for p in self.target_projector.parameters():
            p.requires_grad = False
        self._ema_plan = None

    @staticmethod
    def _copy_weights(source: nn.Module, target: nn.Module) -> None:
        """Parameters only -- buffers (BatchNorm statistics) are not copied, as in ref:src/models/byol.py:57-60."""
        with torch.no_grad():
            for s, t in zip(source.parameters(), target.parameters()):
                t.copy_(s)

    def _ema_pairs(self):
        online = list(self.online_encoder.parameters()) + list(self.online_projector.parameters())
        target = list(self.target_encoder.parameters()) + list(self.target_projector.parameters())
        return online, target

    @torch.no_grad()
    def _update_target_network(self) -> None:
        """target = ema_decay * target + (1 - ema_decay) * online for encoder + projector parameters (496 tensors for
        WavLM-large), one kernel launch.  BatchNorm buffers are not averaged (ref:src/models/byol.py:62-73)."""
        if self._ema_plan is None:
            online, target = self._ema_pairs()
            self._ema_plan = ops.EmaPlan([p.data for p in online], [p.data for p in target])
        self._ema_plan.step(self.ema_decay)

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() re-allocate parameters: rebuild the chunk table lazily
        self._ema_plan = None
        return super()._apply(fn, *args, **kwargs)

    def _pool(self, hidden: torch.Tensor) -> torch.Tensor:
        if hidden.dim() == 3 and self.pooling == "mean":
            return hidden.mean(dim=1)
        return hidden

    def forward(self, clean_input_values: torch.Tensor, noisy_input_values: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        online_emb = self._pool(self.online_encoder(clean_input_values))
        online_pred = self.online_predictor(self.online_projector(online_emb))
        with torch.no_grad():
            target_emb = self._pool(self.target_encoder(noisy_input_values))
            target_proj = self.target_projector(target_emb)
        return online_pred, target_proj

    def get_encoder(self) -> nn.Module:
        return self.online_encoder


def log_loss_flags(flags: int) -> None:
    """The reference's two diagnostics (ref:src/models/byol.py:109-122) from the loss kernel's flag word."""
    if flags & 0x3:
        logger.error("NaN detected in tensors before normalization!")
    if flags & 0xC:
        logger.error("NaN detected in tensors after normalization!")


def byol_loss(online_pred: torch.Tensor, target_proj: torch.Tensor, check_finite: bool = False) -> torch.Tensor:
    """2 - 2 * mean_b clamp(<normalize(p + 1e-10), normalize(z + 1e-10)>, -1, 1)  (ref:src/models/byol.py:104-129).
    ``check_finite=True`` reproduces the reference's two NaN diagnostics on the INPUTS (before / after normalisation):
    the same launch writes a flag word, and reading it is one host synchronisation (the reference pays four)."""
    if not check_finite:
        return ops.byol_loss(online_pred, target_proj)
    loss, flags = ops.byol_loss_with_flags(online_pred, target_proj)
    log_loss_flags(int(flags.item()))
    return loss
