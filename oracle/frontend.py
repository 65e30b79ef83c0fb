"""Oracle: WavLM convolutional feature encoder, fp32 on the CPU (TEST INFRASTRUCTURE ONLY).

The arithmetic lives in a third-party dependency of the reference: HuggingFace
``transformers`` (unpinned in ref:requirements.txt:3; 5.5.0 in this image), reached through
``AutoModel.from_pretrained`` at ref:src/models/encoder.py:14 and ``self.model(input_values)``
at ref:src/models/encoder.py:25.  This file restates

* ``WavLMFeatureEncoder.forward``      hf:models/wavlm/modeling_wavlm.py:779-789
* ``WavLMLayerNormConvLayer.forward``  hf:models/wavlm/modeling_wavlm.py:720-727  ("layer", wavlm-large)
* ``WavLMGroupNormConvLayer.forward``  hf:models/wavlm/modeling_wavlm.py:747-751  ("group", layer 0 of base)
* ``WavLMNoLayerNormConvLayer.forward``hf:models/wavlm/modeling_wavlm.py:697-700  ("group", layers 1-6)

with plain torch functional ops; ``tests/test_oracle_golden.py`` checks it against the
installed ``transformers`` classes and against fixtures produced through the reference's
``WavLMEncoder`` wrapper.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)


def conv_out_lengths(n_samples: int):
    out, t = [], int(n_samples)
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        t = (t - k) // s + 1
        out.append(t)
    return out


def conv_frontend(x: torch.Tensor, layers, norm_mode: str = "layer", return_all: bool = False):
    """x [B,L] (or [B,1,L], squeezed as ref:src/models/encoder.py:20-21) fp32 -> [B,512,T] fp32.

    ``layers``: 7 dicts {"conv": [512,Cin,k], "gamma": [512]|None, "beta": [512]|None}.
    """
    if x.dim() == 3:
        x = x.squeeze(1)
    h = x[:, None]                                                    # hf:...:780
    outs = []
    for i, layer in enumerate(layers):
        w = torch.as_tensor(layer["conv"], dtype=torch.float32)
        h = F.conv1d(h, w, bias=None, stride=CONV_STRIDE[i])          # conv_bias=False for both variants
        if norm_mode == "layer":
            g = torch.as_tensor(layer["gamma"], dtype=torch.float32)
            b = torch.as_tensor(layer["beta"], dtype=torch.float32)
            h = h.transpose(-2, -1)                                   # hf:...:723
            h = F.layer_norm(h, (h.shape[-1],), g, b, eps=1e-5)       # hf:...:724
            h = h.transpose(-2, -1)                                   # hf:...:725
        elif norm_mode == "group" and i == 0:
            g = torch.as_tensor(layer["gamma"], dtype=torch.float32)
            b = torch.as_tensor(layer["beta"], dtype=torch.float32)
            h = F.group_norm(h, h.shape[1], g, b, eps=1e-5)           # hf:...:749 (512 groups of 1 channel)
        elif norm_mode != "group":
            raise ValueError(f"norm_mode must be 'layer' or 'group', got {norm_mode!r}")
        h = F.gelu(h)                                                 # exact erf GELU (ACT2FN['gelu'])
        if return_all:
            outs.append(h)
    return outs if return_all else h
