"""Deterministic synthetic inputs for tests, golden fixtures and bench.py.

Everything here is drawn from ``numpy.random.RandomState`` (MT19937 + the legacy
Box-Muller ``standard_normal``), whose stream is frozen across NumPy versions, so the
same seed gives bit-identical float32 arrays in the build container and on the GPU box.

Recipe: clean = 0.1*randn*U(0.2,1) per row, with a zero-padded tail on
25 % of rows (mimics the crop/pad in ref:src/utils/audio_utils.py:38-49); noise =
0.3*randn; SNR index uniform over ``snr_range``.
"""
from __future__ import annotations

import numpy as np

DEFAULT_SNR_RANGE = (2, 5, 10, 15, 20)  # ref:config/default_wavlm-large_byol.yaml:27

# WavLM conv feature encoder geometry (hf:models/wavlm/configuration_wavlm.py defaults)
CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
CONV_DIM = 512


def conv_out_lengths(n_samples: int) -> list[int]:
    """Per-layer output length T_i = (T_{i-1} - k_i)//s_i + 1 (hf:...modeling_wavlm.py:1000-1014)."""
    out, t = [], int(n_samples)
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        t = (t - k) // s + 1
        out.append(t)
    return out


def waveforms(batch: int, n_samples: int, seed: int = 1234, n_noise: int | None = None,
              snr_range=DEFAULT_SNR_RANGE, pad_fraction: float = 0.25):
    """Returns (clean [B,L] f32, noise [B,Ln] f32, snr_idx [B] i32, snr_table [n] f64)."""
    rs = np.random.RandomState(seed)
    n_noise = n_samples if n_noise is None else n_noise
    gain = rs.uniform(0.2, 1.0, size=(batch, 1))
    clean = (0.1 * rs.standard_normal((batch, n_samples)) * gain).astype(np.float32)
    pad_rows = rs.uniform(size=batch) < pad_fraction
    pad_len = rs.randint(n_samples // 8, n_samples // 2 + 1, size=batch)
    for b in range(batch):
        if pad_rows[b]:
            clean[b, n_samples - pad_len[b]:] = 0.0
    noise = (0.3 * rs.standard_normal((batch, n_noise))).astype(np.float32)
    snr_idx = rs.randint(0, len(snr_range), size=batch).astype(np.int32)
    snr_table = np.asarray(snr_range, dtype=np.float64)
    return clean, noise, snr_idx, snr_table


def frontend_weights(norm_mode: str = "layer", seed: int = 0):
    """Random conv-frontend parameters with the shapes of hf WavLMFeatureEncoder.

    Returns a list of 7 dicts {"conv": [512,Cin,k] f32, "gamma": [512] f32 | None,
    "beta": [512] f32 | None}.  Conv weights follow Kaiming-normal (the HF init for
    nn.Conv1d in WavLM, hf:...modeling_wavlm.py `_init_weights`), norm affine parameters
    are perturbed around (1, 0) so that parity tests exercise them.
    """
    rs = np.random.RandomState(seed)
    layers = []
    cin = 1
    for i, k in enumerate(CONV_KERNEL):
        std = np.sqrt(2.0 / (cin * k))
        w = (std * rs.standard_normal((CONV_DIM, cin, k))).astype(np.float32)
        has_norm = norm_mode == "layer" or (norm_mode == "group" and i == 0)
        gamma = (1.0 + 0.1 * rs.standard_normal(CONV_DIM)).astype(np.float32) if has_norm else None
        beta = (0.1 * rs.standard_normal(CONV_DIM)).astype(np.float32) if has_norm else None
        layers.append({"conv": w, "gamma": gamma, "beta": beta})
        cin = CONV_DIM
    return layers


def embeddings(batch: int, dim: int = 1024, seed: int = 7):
    """(online_pred, target_proj) pair [B,D] f32 with correlated rows (cos sim spread over (-1,1))."""
    rs = np.random.RandomState(seed)
    p = rs.standard_normal((batch, dim)).astype(np.float32)
    mix = rs.uniform(-1.0, 1.0, size=(batch, 1)).astype(np.float32)
    z = (mix * p + (1 - np.abs(mix)) * rs.standard_normal((batch, dim))).astype(np.float32)
    return p, z


def optim_step_grad(seed: int, k: int, i: int, shape, scale: float) -> np.ndarray:
    """Synthetic gradient of tensor ``i`` at step ``k`` of the optimizer fixture (same generator as
    tests/golden/make_golden.py::optim_step_grad, so the fixture does not have to store the gradients)."""
    return (scale * np.random.RandomState(seed * 100000 + k * 1000 + i).standard_normal(shape)).astype(np.float32)
