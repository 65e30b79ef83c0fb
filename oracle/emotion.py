"""Oracle: the consumer side of the hot path on the emotion fine-tune step (TEST INFRASTRUCTURE ONLY).

* ``attentive_statistics_pooling``  restates AttentiveStatisticsPooling.forward, ref:src/models/pool.py:37-58
  (per-utterance loop, exactly the reference's ops in the reference's order)
* ``compute_length_from_mask``      ref:src/models/pool.py:11-19
* ``ccc_loss``                      ref:src/train/dimentional_emotions.py:427-450 (per-dimension loop)

Pinned by ``tests/golden/emotion.npz`` (outputs and autograd gradients of the reference's own classes,
tests/golden/make_golden.py::gen_emotion).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def compute_length_from_mask(mask: torch.Tensor):
    wav_lens = torch.sum(mask, dim=1)                                                   # :16
    feat_lens = torch.div(wav_lens - 1, 16000 * 0.02, rounding_mode="floor") + 1        # :17
    return feat_lens.int().tolist()                                                     # :18


def attentive_statistics_pooling(xs: torch.Tensor, mask: torch.Tensor, sap_weight: torch.Tensor,
                                 sap_bias: torch.Tensor, attention: torch.Tensor) -> torch.Tensor:
    """xs [B,T,D], mask [B,L], sap_linear = (weight [D,D], bias [D]), attention [D,1] -> [B,2D]."""
    feat_lens = compute_length_from_mask(mask)                                          # :44
    pooled = []
    for x, feat_len in zip(xs, feat_lens):                                              # :46
        x = x[:feat_len].unsqueeze(0)                                                   # :47
        h = torch.tanh(F.linear(x, sap_weight, sap_bias))                               # :48
        w = torch.matmul(h, attention).squeeze(dim=2)                                   # :49
        w = F.softmax(w, dim=1).view(x.size(0), x.size(1), 1)                           # :50
        mu = torch.sum(x * w, dim=1)                                                    # :55
        rh = torch.sqrt((torch.sum((x ** 2) * w, dim=1) - mu ** 2).clamp(min=1e-5))     # :56
        pooled.append(torch.cat((mu, rh), 1).squeeze(0))                                # :57
    return torch.stack(pooled)                                                          # :59


def ccc_loss(predictions: torch.Tensor, targets: torch.Tensor):
    batch_size = predictions.size(0)
    loss = 0.0
    if batch_size > 1:
        for i in range(predictions.size(1)):
            pred, target = predictions[:, i], targets[:, i]
            mean_pred, mean_target = torch.mean(pred), torch.mean(target)
            var_pred, var_target = torch.var(pred, unbiased=False), torch.var(target, unbiased=False)
            covar = torch.mean((pred - mean_pred) * (target - mean_target))
            ccc = 2 * covar / (var_pred + var_target + (mean_pred - mean_target) ** 2 + 1e-10)
            loss += 1 - ccc
    return loss / predictions.size(1)
