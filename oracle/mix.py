"""Oracle: SNR mixing + peak normalisation + z-normalisation (TEST INFRASTRUCTURE ONLY).

Restates, op for op and in the same fp32 order, what the reference does per utterance on
a DataLoader worker:

* ``add_noise_to_speech``            ref:src/data/augment.py:4-66
* peak normalisation                 ref:src/data/noisy_speech_dataset.py:88-116
* HF ``zero_mean_unit_var_norm``     hf:models/wav2vec2/feature_extraction_wav2vec2.py:78-97
  (called through ``self.feature_extractor(x.squeeze().numpy(), ...)`` at
  ref:src/data/noisy_speech_dataset.py:120-129 and ref:src/data/emotion_dataset.py:198-203)

The reference signals failure by returning ``None`` / retrying; here every exit is given a
status code (the same codes the CUDA kernel writes per row) so the two can be compared.
"""
from __future__ import annotations

import numpy as np
import torch

STATUS_OK = 0
STATUS_SPEECH_NAN = 1        # ref:src/data/augment.py:7-9
STATUS_NOISE_NAN = 2         # ref:src/data/augment.py:11-13
STATUS_SPEECH_POWER = 3      # ref:src/data/augment.py:30-32
STATUS_NOISE_POWER = 4       # ref:src/data/augment.py:34-36
STATUS_SCALE_INVALID = 5     # ref:src/data/augment.py:45-47
STATUS_SCALE_LARGE = 6       # ref:src/data/augment.py:49-51
STATUS_SCALED_NOISE_NAN = 7  # ref:src/data/augment.py:56-58
STATUS_NOISY_NAN = 8         # ref:src/data/augment.py:62-64
STATUS_CLEAN_PEAK = 9        # ref:src/data/noisy_speech_dataset.py:95-97
STATUS_NOISY_PEAK = 10       # ref:src/data/noisy_speech_dataset.py:99-101
STATUS_CLEAN_NORM_NAN = 11   # ref:src/data/noisy_speech_dataset.py:107-109
STATUS_NOISY_NORM_NAN = 12   # ref:src/data/noisy_speech_dataset.py:111-113
STATUS_CLEAN_ZNORM_NAN = 13  # ref:src/data/noisy_speech_dataset.py:132-134
STATUS_NOISY_ZNORM_NAN = 14  # ref:src/data/noisy_speech_dataset.py:136-138

STATUS_NAMES = {
    0: "ok", 1: "speech_nan", 2: "noise_nan", 3: "speech_power_too_small",
    4: "noise_power_too_small", 5: "scale_invalid", 6: "scale_too_large",
    7: "scaled_noise_nan", 8: "noisy_nan", 9: "clean_peak_too_small",
    10: "noisy_peak_too_small", 11: "clean_norm_nan", 12: "noisy_norm_nan",
    13: "clean_znorm_nan", 14: "noisy_znorm_nan",
}


def _add_noise_status(speech: torch.Tensor, noise: torch.Tensor, snr_db: float):
    """ref:src/data/augment.py:4-66 with the ``return None`` exits numbered. -> (noisy|None, status)"""
    if torch.isnan(speech).any():                                   # :7
        return None, STATUS_SPEECH_NAN
    if torch.isnan(noise).any():                                    # :11
        return None, STATUS_NOISE_NAN
    if noise.shape[1] > speech.shape[1]:                            # :16-17 truncate
        noise = noise[:, :speech.shape[1]]
    elif noise.shape[1] < speech.shape[1]:                          # :18-21 tile
        repetitions = (speech.shape[1] // noise.shape[1]) + 1
        noise = noise.repeat(1, repetitions)[:, :speech.shape[1]]
    speech_power = torch.mean(speech ** 2)                          # :24
    noise_power = torch.mean(noise ** 2)                            # :25
    if speech_power < 1e-10:                                        # :30
        return None, STATUS_SPEECH_POWER
    if noise_power < 1e-10:                                         # :34
        return None, STATUS_NOISE_POWER
    snr_linear = 10 ** (snr_db / 10)                                # :39 (Python double)
    noise_scaling = torch.sqrt(speech_power / (noise_power * snr_linear))   # :40 (fp32)
    if torch.isinf(noise_scaling) or torch.isnan(noise_scaling):    # :45
        return None, STATUS_SCALE_INVALID
    if noise_scaling > 1e6:                                         # :49
        return None, STATUS_SCALE_LARGE
    scaled_noise = noise * noise_scaling                            # :54
    if torch.isnan(scaled_noise).any():                             # :56
        return None, STATUS_SCALED_NOISE_NAN
    noisy_speech = speech + scaled_noise                            # :60
    if torch.isnan(noisy_speech).any():                             # :62
        return None, STATUS_NOISY_NAN
    return noisy_speech, STATUS_OK


def add_noise_to_speech(speech: torch.Tensor, noise: torch.Tensor, snr_db: float):
    """Same signature and ``None`` semantics as ref:src/data/augment.py:4."""
    return _add_noise_status(speech, noise, snr_db)[0]


def peak_normalize_pair(clean: torch.Tensor, noisy: torch.Tensor):
    """ref:src/data/noisy_speech_dataset.py:88-116. -> (clean, noisy, status)"""
    clean_max = torch.max(torch.abs(clean))                         # :91
    noisy_max = torch.max(torch.abs(noisy))                         # :92
    if clean_max < 1e-8:                                            # :95
        return None, None, STATUS_CLEAN_PEAK
    if noisy_max < 1e-8:                                            # :99
        return None, None, STATUS_NOISY_PEAK
    clean = clean / (clean_max + 1e-8)                              # :103
    noisy = noisy / (noisy_max + 1e-8)                              # :104
    if torch.isnan(clean).any():                                    # :107
        return None, None, STATUS_CLEAN_NORM_NAN
    if torch.isnan(noisy).any():                                    # :111
        return None, None, STATUS_NOISY_NORM_NAN
    return clean, noisy, STATUS_OK


def zero_mean_unit_var_norm(x: np.ndarray) -> np.ndarray:
    """hf:models/wav2vec2/feature_extraction_wav2vec2.py:95 (attention_mask is None): float32 numpy."""
    x = np.asarray(x, dtype=np.float32)
    return (x - x.mean()) / np.sqrt(x.var() + 1e-7)


def _znorm_tensor(x: torch.Tensor) -> torch.Tensor:
    """``feature_extractor(x.squeeze().numpy(), return_tensors='pt').input_values`` -> [1, L]."""
    return torch.from_numpy(zero_mean_unit_var_norm(x.squeeze().numpy()))[None, :]


def mix_normalize_item(clean: torch.Tensor, noise: torch.Tensor, snr_db: float, peak_norm: bool = True):
    """One utterance through the worker-side chain.  clean [1,L], noise [1,Ln] fp32.

    peak_norm=True : ref:src/data/noisy_speech_dataset.py:54-148 (BYOL pre-training) ->
                     (clean_input_values [1,L] | None, noisy_input_values [1,L] | None, status)
    peak_norm=False: ref:src/data/emotion_dataset.py:177-203 (fine-tuning: no peak-norm; a failed
                     mix keeps the clean waveform, :193-194) -> (None, input_values [1,L], status)
    """
    noisy, status = _add_noise_status(clean, noise, snr_db)
    if not peak_norm:
        wave = clean if noisy is None else noisy
        return None, _znorm_tensor(wave), status
    if noisy is None:
        return None, None, status
    clean_n, noisy_n, status = peak_normalize_pair(clean, noisy)
    if status != STATUS_OK:
        return None, None, status
    clean_z = _znorm_tensor(clean_n)
    noisy_z = _znorm_tensor(noisy_n)
    if torch.isnan(clean_z).any():
        return None, None, STATUS_CLEAN_ZNORM_NAN
    if torch.isnan(noisy_z).any():
        return None, None, STATUS_NOISY_ZNORM_NAN
    return clean_z, noisy_z, STATUS_OK


def mix_normalize_batch(clean, noise, snr_idx, snr_table, peak_norm: bool = True):
    """Row-by-row application of ``mix_normalize_item`` (the DataLoader collate, ref:train_byol.py:49-50).

    clean [B,L] f32, noise [B,Ln] f32, snr_idx [B] int, snr_table [n] (dB values, as in the YAML
    ``data.snr_range``).  Rows whose status != 0 are zero-filled (the reference would re-draw).
    Returns (clean_out [B,L] | None, noisy_out [B,L], status [B] int32).
    """
    clean = torch.as_tensor(clean, dtype=torch.float32)
    noise = torch.as_tensor(noise, dtype=torch.float32)
    B, L = clean.shape
    clean_out = torch.zeros(B, L) if peak_norm else None
    noisy_out = torch.zeros(B, L)
    status = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        snr_db = snr_table[int(snr_idx[b])]
        snr_db = snr_db.item() if hasattr(snr_db, "item") else snr_db
        c, n, st = mix_normalize_item(clean[b:b + 1], noise[b:b + 1], snr_db, peak_norm)
        status[b] = st
        if c is not None:
            clean_out[b] = c[0]
        if n is not None:
            noisy_out[b] = n[0]
    return clean_out, noisy_out, status


def mix_batch_attempts(clean, noise, snr_idx, snr_table, max_attempts: int = 5, substitute: bool = True):
    """The attempt loop of ``NoiseRobustSpeechDataset.__getitem__`` (ref:src/data/noisy_speech_dataset.py:55-149) for a
    batch, with the reference's random re-draws replaced by EXPLICIT donors so that it can be compared bit for bit:

    * ``for attempt in range(max_attempts)`` (:58): attempt a of row b mixes clean[b] with the noise crop AND the SNR
      draw of row ``(b + a) % B`` -- attempt 0 is the row's own draw; every later attempt "loads another random noise"
      (:69-70) and "selects a random SNR" (:78) again, here the following rows' draws;
    * the first attempt that passes every check (:81-138) is the item (:140-144); ``snr`` is the table value it was
      mixed at;
    * a row that fails all attempts is not emitted by the reference (:146-149 raises after moving on, :60-66); with
      ``substitute`` it takes the outputs and the SNR of the nearest following row that passed, otherwise it stays
      zero-filled with its last status.

    A single-row batch has no donors: one attempt.  Returns (clean_out [B,L], noisy_out [B,L], status [B] int32 -- 0 or
    the last attempt's exit --, snr_idx_used [B] int32, n_rejected).
    """
    clean = torch.as_tensor(clean, dtype=torch.float32)
    noise = torch.as_tensor(noise, dtype=torch.float32)
    snr_idx = np.asarray(snr_idx, dtype=np.int32)
    B, L = clean.shape
    clean_out, noisy_out = torch.zeros(B, L), torch.zeros(B, L)
    status = torch.zeros(B, dtype=torch.int32)
    used = snr_idx.copy()
    for b in range(B):
        for attempt in range(max_attempts if B > 1 else 1):                       # :58
            donor = (b + attempt) % B
            snr_db = snr_table[int(snr_idx[donor])]
            snr_db = snr_db.item() if hasattr(snr_db, "item") else snr_db
            c, n, st = mix_normalize_item(clean[b:b + 1], noise[donor:donor + 1], snr_db, True)
            status[b], used[b] = st, snr_idx[donor]
            if st == STATUS_OK:                                                   # :140-144
                clean_out[b], noisy_out[b] = c[0], n[0]
                break
    good = [int(status[b]) == STATUS_OK for b in range(B)]
    n_rejected = B - sum(good)
    if substitute and B > 1:
        src_c, src_n, src_used = clean_out.clone(), noisy_out.clone(), used.copy()
        for b in range(B):
            if good[b]:
                continue
            for d in range(1, B):                                                 # "move on to the next item", :60-66
                r = (b + d) % B
                if good[r]:
                    clean_out[b], noisy_out[b], used[b] = src_c[r], src_n[r], src_used[r]
                    break
    return clean_out, noisy_out, status, used, n_rejected
