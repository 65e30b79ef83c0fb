"""CPU oracle for the BYOL noisy-view hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference
(sunYtokki/Noise-Robust-Speech-Embedding) for the one path this repository accelerates:
SNR mix -> peak-normalise -> z-normalise -> WavLM conv feature encoder -> BYOL cosine
loss -> EMA target update.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product package (``nrse_b200``) never does: it
fails loudly when the CUDA library is missing instead of falling back to this code.

Pinning: the reference ships no golden vectors or tests for this path,
so the oracle is pinned against outputs of the *reference itself*, imported unmodified
from /root/reference in the build container by ``tests/golden/make_golden.py`` and
committed as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every oracle function against them.  The conv feature encoder and the z-normalisation
live in a third-party dependency of the reference (HuggingFace ``transformers``,
unpinned in ref:requirements.txt:3; 5.5.0 in this image); ``oracle.frontend`` restates
them with plain torch ops and is additionally checked against the installed
``transformers`` classes, which are present on the GPU box too.
"""
from .mix import (  # noqa: F401
    STATUS_OK,
    STATUS_NAMES,
    add_noise_to_speech,
    mix_batch_attempts,
    mix_normalize_batch,
    mix_normalize_item,
    peak_normalize_pair,
    zero_mean_unit_var_norm,
)
from .byol import byol_loss, ema_update  # noqa: F401
from .frontend import conv_frontend, conv_out_lengths  # noqa: F401
from .emotion import attentive_statistics_pooling, ccc_loss, compute_length_from_mask  # noqa: F401
from .optim import adamw_step, clip_adamw_ema_step, clip_grad_norm  # noqa: F401
from .evaluate import embedding_similarity, validation_metrics  # noqa: F401
