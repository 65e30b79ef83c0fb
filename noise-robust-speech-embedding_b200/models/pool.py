"""``AttentiveStatisticsPooling`` with the reference's surface (ref:src/models/pool.py:7-58): parameters ``sap_linear``
(Linear D->D) and ``attention`` ([D,1], N(0,1) init) -- the checkpoint contract -- and ``forward(xs, mask)`` returning
``[B, 2D]``.  The per-utterance Python loop of the reference (slice, linear, tanh, matmul, softmax, two weighted sums,
~12 launches per utterance) is one library GEMM over all frames plus the two launches of ``ops.asp_pool``."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class Pooling(nn.Module):
    def compute_length_from_mask(self, mask: torch.Tensor) -> torch.Tensor:
        """Frames per utterance for a sample-level mask ``[B, L]`` at 16 kHz / 20 ms frame shift
        (ref:src/models/pool.py:11-19).  Returns an int tensor on the mask's device (the reference returns a Python
        list, which costs a device-to-host sync per batch)."""
        wav_lens = torch.sum(mask, dim=1)
        feat_lens = torch.div(wav_lens - 1, 16000 * 0.02, rounding_mode="floor") + 1
        return feat_lens.int()

    def forward(self, x, mask):
        raise NotImplementedError


class AttentiveStatisticsPooling(Pooling):
    """Attentive Statistics Pooling (arXiv:1803.10963), ref:src/models/pool.py:24-58."""

    def __init__(self, input_size: int):
        super().__init__()
        self._indim = input_size
        self.sap_linear = nn.Linear(input_size, input_size)
        self.attention = nn.Parameter(torch.FloatTensor(input_size, 1))
        torch.nn.init.normal_(self.attention, mean=0, std=1)

    def forward(self, xs: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """xs: [B, T, D]; mask: [B, L] (sample-level, as the reference's callers pass) -> [B, 2D]."""
        feat_lens = self.compute_length_from_mask(mask).clamp(max=xs.shape[1])  # x[:feat_len] never exceeds T (:47)
        xs = xs.float()
        hl = self.sap_linear(xs)  # one GEMM over all B*T frames; tanh and everything after it is fused in the op
        return ops.asp_pool(xs, hl, self.attention, feat_lens)
