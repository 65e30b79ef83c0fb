"""``WavLMEncoder`` with the reference's surface (ref:src/models/encoder.py:5-32): ``.model`` is the HF ``WavLMModel``
(so ``.model.named_parameters()`` / ``.model.config.hidden_size`` keep their meaning), ``.output_dim``, and
``forward(input_values[, attention_mask])`` accepting ``[B,L]`` or ``[B,1,L]`` and returning ``last_hidden_state``.
The only change: ``.model.feature_extractor`` runs the B200 conv-frontend kernels."""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.nn as nn
from transformers import AutoModel, WavLMConfig, WavLMModel

from .featproj import B200FeatureProjection
from .frontend import B200FeatureEncoder
from .posconv import B200PositionalConvEmbedding


def wavlm_large_config(**overrides) -> WavLMConfig:
    """microsoft/wavlm-large hyper-parameters for offline, random-init construction."""
    kw = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
              feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False, conv_dim=(512,) * 7,
              conv_kernel=(10, 3, 3, 3, 3, 2, 2), conv_stride=(5, 2, 2, 2, 2, 2, 2), feat_extract_activation="gelu",
              mask_time_prob=0.075, layerdrop=0.1, num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16)
    kw.update(overrides)
    return WavLMConfig(**kw)


def _upload_mask(mask_np, device) -> torch.Tensor:
    """numpy bool mask -> device tensor through pinned memory, non-blocking (``torch.tensor(np, device=cuda)`` is a
    blocking copy from pageable memory: the host waits for everything queued on the stream before it)."""
    host = torch.from_numpy(mask_np)
    if device.type != "cuda":
        return host.to(device)
    return host.pin_memory().to(device, non_blocking=True)


def install_sync_free_spec_augment(model: nn.Module) -> nn.Module:
    """Replace ``WavLMModel._mask_hidden_states`` (hf:models/wavlm/modeling_wavlm.py, SpecAugment on the projected
    features, active in training mode with wavlm-large's ``mask_time_prob`` = 0.075) by an equivalent that never
    synchronises the host with the device.  The stock method synchronises twice per forward: the mask is uploaded with a
    blocking ``torch.tensor(..., device=...)`` and applied with boolean-mask assignment (``index_put_`` -> ``nonzero``
    -> device-to-host read).  Here the same ``_compute_mask_indices`` call (same numpy RNG stream, same arguments)
    produces the mask, it travels through pinned memory, and ``torch.where`` applies it: same values, same gradients."""
    from transformers.models.wavlm import modeling_wavlm as hf

    def _mask_hidden_states(hidden_states, mask_time_indices=None, attention_mask=None):
        cfg = model.config
        if not getattr(cfg, "apply_spec_augment", True):
            return hidden_states
        batch_size, sequence_length, hidden_size = hidden_states.size()
        embed = model.masked_spec_embed.to(hidden_states.dtype)
        if mask_time_indices is not None:
            hidden_states = torch.where(mask_time_indices.bool()[:, :, None], embed, hidden_states)
        elif cfg.mask_time_prob > 0 and model.training:
            mask = hf._compute_mask_indices((batch_size, sequence_length), mask_prob=cfg.mask_time_prob,
                                            mask_length=cfg.mask_time_length, attention_mask=attention_mask,
                                            min_masks=cfg.mask_time_min_masks)
            mask = _upload_mask(mask, hidden_states.device)
            hidden_states = torch.where(mask[:, :, None], embed, hidden_states)
        if cfg.mask_feature_prob > 0 and model.training:
            mask = hf._compute_mask_indices((batch_size, hidden_size), mask_prob=cfg.mask_feature_prob,
                                            mask_length=cfg.mask_feature_length,
                                            min_masks=getattr(cfg, "mask_feature_min_masks", 0))
            mask = _upload_mask(mask, hidden_states.device)
            hidden_states = hidden_states.masked_fill(mask[:, None, :], 0)
        return hidden_states

    model._mask_hidden_states = _mask_hidden_states
    return model


def install_b200_frontend(model: nn.Module) -> nn.Module:
    """Swap ``model.feature_extractor`` (HF WavLMFeatureEncoder) for the B200 implementation, in place, and -- for the
    wavlm-large geometry -- ``model.feature_projection`` (512 -> 1024) and
    ``model.encoder.pos_conv_embed`` (1024 channels, 16 groups, 128 taps, 8f-4) too; other sizes keep HF's modules."""
    B200FeatureEncoder.convert(model.feature_extractor)
    if B200FeatureProjection.supports(getattr(model, "feature_projection", None)):
        B200FeatureProjection.convert(model.feature_projection)
    pos = getattr(getattr(model, "encoder", None), "pos_conv_embed", None)
    if B200PositionalConvEmbedding.supports(pos):   # wavlm-large's grouped positional convolution
        B200PositionalConvEmbedding.convert(pos)
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
            install_sync_free_spec_augment(self.model)
        elif frontend != "hf":
            raise ValueError("frontend must be 'b200' or 'hf'")
        self.output_dim = self.model.config.hidden_size

    def forward(self, input_values: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if input_values.dim() == 3:  # [B, 1, L] -> [B, L]
            input_values = input_values.squeeze(1)
        # the attention mask is accepted and ignored, exactly like the reference (encoder.py:23-25)
        outputs = self.model(input_values)
        return outputs.last_hidden_state if hasattr(outputs, "last_hidden_state") else outputs
