"""Oracle: the validation passes of the BYOL step (TEST INFRASTRUCTURE ONLY).

* ``embedding_similarity``  restates ref:evaluate_byol.py:12-66: per-SNR mean cosine similarity of clean / noisy embeddings
* ``validation_metrics``    restates ref:evaluate_byol.py:69-123: mean ``byol_loss`` over the batches + the metrics dict

Both take the model's outputs (embeddings / (online_pred, target_proj) pairs per batch) instead of a model: what is restated
is the arithmetic after the encoder.  Pinned by tests/golden/evaluate_byol.npz, which the reference's own functions wrote.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .byol import byol_loss


def embedding_similarity(batches, snr_range):
    """batches: iterable of (clean_emb [B,H], noisy_emb [B,H], snr [B] ints).  -> {snr: mean similarity (0 if none)}."""
    sims = {snr: [] for snr in snr_range}                                   # :33
    for clean_emb, noisy_emb, snr in batches:
        c = F.normalize(clean_emb, dim=1)                                    # :51
        n = F.normalize(noisy_emb, dim=1)                                    # :52
        similarity = torch.sum(c * n, dim=1)                                 # :55
        for idx, s in enumerate(torch.as_tensor(snr).tolist()):             # :58-60
            if s in sims:
                sims[s].append(similarity[idx].item())
    return {s: sum(v) / len(v) if len(v) > 0 else 0 for s, v in sims.items()}   # :63-64


def validation_metrics(pairs, similarities):
    """pairs: iterable of (online_pred, target_proj) per batch.  -> (val_loss, metrics) as ref:evaluate_byol.py:94-123."""
    total, n = 0.0, 0
    for online_pred, target_proj in pairs:
        total += byol_loss(online_pred, target_proj).item()                  # :104-107
        n += 1
    val_loss = total / n if n > 0 else float("inf")                          # :111
    avg = sum(similarities.values()) / len(similarities) if similarities else 0.0   # :114
    return val_loss, {"val_loss": val_loss, "val_avg_similarity": avg, "val_similarities": similarities}
