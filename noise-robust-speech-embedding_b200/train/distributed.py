"""Data parallelism for the BYOL step: one process per GPU, the batch sharded across ranks, ONE gradient all-reduce
per step through ``DistributedDataParallel`` (NCCL over NVLink on the B200 box, gloo in the CPU tests).

The hot-path kernels need no collective: utterances are independent through mix, conv frontend and the per-row
cosine; the EMA update is local because every rank holds identical online weights after the all-reduced optimizer
step and therefore computes identical targets.  Only trainable (online) parameters take part in the
all-reduce -- the target network has ``requires_grad=False`` and DDP skips it.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment; a no-op single process when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend)
    return rank, world, local_rank


def wrap_data_parallel(model: torch.nn.Module, device: torch.device, bucket_cap_mb: int = 100,
                       find_unused_parameters: bool = True):
    """DDP over the trainable (online) parameters; BatchNorm buffers are NOT broadcast each step, so every rank keeps
    the per-rank batch statistics the single-GPU reference would have on the same per-rank inputs.
    ``find_unused_parameters`` defaults to True because WavLM's LayerDrop (0.1 in wavlm-large) skips whole transformer
    layers at random, per rank, in training mode: those parameters get no gradient in that iteration."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    ids = [device.index] if device.type == "cuda" else None
    return DistributedDataParallel(model, device_ids=ids, broadcast_buffers=False, bucket_cap_mb=bucket_cap_mb,
                                   gradient_as_bucket_view=True, find_unused_parameters=find_unused_parameters)
