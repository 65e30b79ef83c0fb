"""``EmotionClassifier`` with the reference's surface (ref:src/models/emotion.py:8-133): sub-modules ``encoder``,
``pooling`` (AttentiveStatisticsPooling), ``shared_fc``, ``categorical_fc``, ``categorical_out``, ``dimensional_fc``,
``dimensional_out`` -- same names, shapes and state-dict keys -- ``forward(x, attention_mask, task)`` returning
``(categorical_logits, dimensional_values)``, and the freeze / gradual-unfreeze helpers whose substring matching the
B200 feature encoder keeps valid (its parameters are still called ``feature_extractor.conv_layers.{i}...``).

The encoder is the consumer-side of the hot path (config 4 of BASELINE.json): its conv feature encoder runs the B200
kernels, the pooling is the batched ``ops.asp_pool``; the MLP heads are stock torch."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .pool import AttentiveStatisticsPooling


class EmotionClassifier(nn.Module):
    def __init__(self, encoder: nn.Module, hidden_dim: int = 1024, dropout: float = 0.5, num_emotions: int = 8):
        super().__init__()
        self.encoder = encoder
        self.input_dim = encoder.output_dim
        self.hidden_dim = hidden_dim
        self.pooling = AttentiveStatisticsPooling(self.input_dim)
        pooled_dim = self.input_dim * 2  # mean | std

        def block(n_in):
            return nn.Sequential(nn.Linear(n_in, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(), nn.Dropout(dropout))
        self.shared_fc = block(pooled_dim)
        self.categorical_fc = block(hidden_dim)
        self.categorical_out = nn.Linear(hidden_dim, num_emotions)
        self.dimensional_fc = block(hidden_dim)
        self.dimensional_out = nn.Linear(hidden_dim, 3)  # arousal, valence, dominance

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                task: str = "both") -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        encoder_outputs = self.encoder(x, attention_mask=attention_mask)
        if attention_mask is None:  # ref:src/models/emotion.py:75-76 -- a [B, T] mask of ones
            attention_mask = torch.ones(x.shape[0], encoder_outputs.shape[1], device=x.device)
        if hasattr(self, "pooling"):
            features = self.pooling(encoder_outputs, attention_mask)
        else:
            features = torch.mean(encoder_outputs, dim=1)
        shared = self.shared_fc(features)
        categorical_logits = dimensional_values = None
        if task in ("categorical", "both"):
            categorical_logits = self.categorical_out(self.categorical_fc(shared))
        if task in ("dimensional", "both"):
            dimensional_values = self.dimensional_out(self.dimensional_fc(shared))
        return categorical_logits, dimensional_values

    def freeze_encoder(self) -> None:
        for p in self.encoder.parameters():
            p.requires_grad = False

    def unfreeze_encoder(self) -> None:
        for p in self.encoder.parameters():
            p.requires_grad = True

    def unfreeze_encoder_gradually(self, layers_to_unfreeze: Sequence[int]) -> None:
        """Freeze everything in ``encoder.model``, then unfreeze parameters whose name contains ``layer.{i}`` or
        ``layers.{i}`` (ref:src/models/emotion.py:114-129) -- a substring match, so index 1 also matches 10..19 and
        ``feature_extractor.conv_layers.{i}``; kept as is."""
        for _, p in self.encoder.model.named_parameters():
            p.requires_grad = False
        for idx in layers_to_unfreeze:
            for name, p in self.encoder.model.named_parameters():
                if f"layer.{idx}" in name or f"layers.{idx}" in name:
                    p.requires_grad = True

    def get_trainable_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)
