"""``add_noise_to_speech`` with the reference's signature and ``None`` semantics (ref:src/data/augment.py:4-66),
computed by the fused CUDA mix kernel (``peak_norm=2`` mode of ``nrse_mix_normalize_f32``): powers, SNR scaling and
``speech + scale * noise`` in one launch instead of ~20 tensor ops and several ``.item()`` syncs per utterance.

The training path does not call this per utterance any more -- ``GpuBatchMixer`` mixes and normalises a whole batch
in one launch -- but the function is kept as the drop-in for notebooks / the emotion dataset call site
(ref:src/data/emotion_dataset.py:190)."""
from __future__ import annotations

from typing import List, Optional, Union

import torch

from .. import ops
from .._lib import NrseError
from ..utils.logging_utils import logger


def add_noise_to_speech(speech: torch.Tensor, noise: torch.Tensor,
                        snr_db: float) -> Union[Optional[torch.Tensor], List[Optional[torch.Tensor]]]:
    """speech [1,L] (or [B,L]), noise [1,Ln] (or [B,Ln]) -> noisy [1,L] or ``None`` (a list for B > 1).

    Inputs are not modified.  CPU tensors are moved to the current CUDA device for the computation and the result is
    returned on the input's device; without a CUDA device this raises (there is no CPU implementation)."""
    if speech.dim() != 2 or noise.dim() != 2 or speech.shape[0] != noise.shape[0]:
        raise NrseError("add_noise_to_speech expects speech [B,L] and noise [B,Ln]")
    src_device = speech.device
    if not speech.is_cuda:
        if not torch.cuda.is_available():
            raise NrseError("add_noise_to_speech needs a CUDA device (nrse_b200 has no CPU fallback)")
        speech, noise = speech.cuda(), noise.cuda()
    B = speech.shape[0]
    idx = torch.zeros(B, dtype=torch.int32, device=speech.device)
    noisy, status = ops.mix_raw(speech, noise.to(speech.device), idx, [float(snr_db)])
    status = status.cpu().tolist()  # the None contract needs the verdict on the host
    out: List[Optional[torch.Tensor]] = []
    for b, st in enumerate(status):
        if st != 0:
            logger.warning("add_noise_to_speech: row %d rejected (%s)", b, ops.mix_status_name(st))
            out.append(None)
        else:
            out.append(noisy[b:b + 1].to(src_device))
    return out[0] if B == 1 else out
