"""``WavLMEncoder`` with the reference's surface (ref:src/models/encoder.py:5-32): ``.model`` is the HF ``WavLMModel``
(so ``.model.named_parameters()`` / ``.model.config.hidden_size`` keep their meaning), ``.output_dim``, and
``forward(input_values[, attention_mask])`` accepting ``[B,L]`` or ``[B,1,L]`` and returning ``last_hidden_state``.
The only change: ``.model.feature_extractor`` runs the B200 conv-frontend kernels."""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.nn as nn
from transformers import AutoModel, WavLMConfig, WavLMModel

from .frontend import B200FeatureEncoder


def wavlm_large_config(**overrides) -> WavLMConfig:
    """microsoft/wavlm-large hyper-parameters for offline, random-init construction."""
    kw = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
              feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False, conv_dim=(512,) * 7,
              conv_kernel=(10, 3, 3, 3, 3, 2, 2), conv_stride=(5, 2, 2, 2, 2, 2, 2), feat_extract_activation="gelu",
              mask_time_prob=0.075, layerdrop=0.1, num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16)
    kw.update(overrides)
    return WavLMConfig(**kw)


def install_b200_frontend(model: nn.Module) -> nn.Module:
    """Swap ``model.feature_extractor`` (HF WavLMFeatureEncoder) for the B200 implementation, in place."""
    B200FeatureEncoder.convert(model.feature_extractor)
    return model


class WavLMEncoder(nn.Module):
    def __init__(self, model_name: Union[str, WavLMConfig], frontend: str = "b200"):
        """model_name: a HF hub id / local path (``AutoModel.from_pretrained``, as the reference does) or a
        ``WavLMConfig`` for random-init construction (no checkpoints are reachable offline)."""
        super().__init__()
        if isinstance(model_name, WavLMConfig):
            self.model = WavLMModel(model_name)
        else:
            self.model = AutoModel.from_pretrained(model_name)
        if frontend == "b200":
            install_b200_frontend(self.model)
        elif frontend != "hf":
            raise ValueError("frontend must be 'b200' or 'hf'")
        self.output_dim = self.model.config.hidden_size

    def forward(self, input_values: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if input_values.dim() == 3:  # [B, 1, L] -> [B, L]
            input_values = input_values.squeeze(1)
        # the attention mask is accepted and ignored, exactly like the reference (encoder.py:23-25)
        outputs = self.model(input_values)
        return outputs.last_hidden_state if hasattr(outputs, "last_hidden_state") else outputs
