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


class GradArena:
    """Gradients of a list of parameter GROUPS in ONE flat fp32 buffer, all-reduced in a few large buckets.

    What ``DistributedDataParallel(gradient_as_bucket_view=True)`` does with hooks, written out so that the step loop
    decides WHEN each group is reduced (and so that gradients which do not come out of autograd -- the fused kernels write
    through raw pointers -- take part):

    * every ``p.grad`` is a view into the arena: addresses never change, so ``FusedAdamWEma`` builds its chunk table once
      (set ``optimizer.keep_grads = True``), ``zero_(group)`` is one memset, and a group is contiguous -- its all-reduce
      is ``ceil(bytes / bucket_bytes)`` NCCL calls on views, no flatten / unflatten copies;
    * ``all_reduce_async(group)`` starts the group's reduction on a side stream -- it overlaps whatever the compute stream
      does next, e.g. the conv-frontend backward -- and ``wait()`` makes the compute stream wait for everything outstanding.
      Default: bucketed ``dist.all_reduce`` calls with ``async_op=True`` (``ReduceOp.AVG`` on NCCL, SUM + divide on gloo).
      Opt-in (``multimem=True``): this repository's own kernel over NVSwitch multicast memory (``csrc/allreduce.cu``: the
      switch sums, every byte crosses a GPU's links once each way) -- correct and tested, but measured no faster than NCCL
      next to the conv backward on this box (profiles/r2_allreduce_overlap.md), hence not the default.

    Groups are reduced in the order the caller asks for: in the BYOL step the transformer / head gradients are complete
    long before the conv frontend's (the frontend is the FIRST layer, its backward runs last), so their buckets travel
    while the frontend backward computes and only the frontend's 16.8 MB are the un-overlappable tail.
    """

    def __init__(self, param_groups, bucket_bytes: int = 256 << 20, process_group=None, multimem=False,
                 multimem_ctas: int = 0):
        """``multimem``: False (default) = bucketed NCCL / gloo all-reduce calls; True / "auto" = reduce through NVSwitch
        multicast memory with this repository's own kernel (csrc/allreduce.cu) when the process group is NCCL on CUDA and
        the platform supports multicast ("auto" falls back to NCCL silently, True raises)."""
        self.groups = [list(g) for g in param_groups]
        params = [p for g in self.groups for p in g]
        if not params:
            raise ValueError("GradArena needs at least one parameter")
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("GradArena: fp32 parameters on one device expected")
        self.process_group = process_group
        self.bucket_elems = max(int(bucket_bytes) // 4, 1)
        # 4-element (16-byte) alignment of every view keeps the vectorised kernels on their fast path
        offsets, off, self.group_ranges = [], 0, []
        for g in self.groups:
            start = off
            for p in g:
                offsets.append(off)
                off += (p.numel() + 3) // 4 * 4
            self.group_ranges.append((start, off))
        self._symm = None
        self.impl = "none (single process)"
        self.flat = None
        if dist.is_initialized() and dist.get_world_size(process_group) > 1:
            self.impl = f"{dist.get_backend(process_group)} all-reduce, {bucket_bytes >> 20} MB buckets"
            if multimem and dev.type == "cuda" and dist.get_backend(process_group) == "nccl":
                try:
                    self._init_multimem(off, dev, multimem_ctas)
                except Exception as e:  # noqa: BLE001 -- any failure of the symmetric-memory set-up means "not available"
                    if multimem is True:
                        raise
                    self._symm = None
                    self.impl += f" (multicast path unavailable: {type(e).__name__})"
        if self.flat is None:
            self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(params, offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
        self._works = []

    def _init_multimem(self, numel: int, dev, max_ctas: int) -> None:
        """The arena as a symmetric allocation bound to one multicast object over all ranks (torch's symmetric-memory
        allocator and rendezvous are the plumbing; the reduction itself is csrc/allreduce.cu)."""
        import torch.distributed._symmetric_memory as symm
        group = self.process_group if self.process_group is not None else dist.group.WORLD
        flat = symm.empty(numel, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(flat, group)
        mc = int(hdl.multicast_ptr)
        if mc == 0:
            raise RuntimeError("no multicast support")
        flat.zero_()
        self._mc_ptr = mc + (flat.data_ptr() - int(hdl.buffer_ptrs[hdl.rank]))
        self._symm, self.flat = hdl, flat
        self._side = torch.cuda.Stream(device=dev)
        self._mm_ctas = int(max_ctas)
        self.impl = "own kernel over NVSwitch multicast memory (multimem.ld_reduce / multimem.st), one launch per group"

    @property
    def numel(self) -> int:
        return self.flat.numel()

    def group_bytes(self, group: int) -> int:
        a, b = self.group_ranges[group]
        return (b - a) * 4

    def zero_(self, group: Optional[int] = None) -> None:
        if group is None:
            self.flat.zero_()
        else:
            a, b = self.group_ranges[group]
            self.flat[a:b].zero_()

    def buckets(self, group: int):
        a, b = self.group_ranges[group]
        return [self.flat[s:min(s + self.bucket_elems, b)] for s in range(a, b, self.bucket_elems)]

    def all_reduce_async(self, group: int) -> int:
        """Enqueue the all-reduce (mean over ranks) of one group; returns the number of collectives issued."""
        if not dist.is_initialized() or dist.get_world_size(self.process_group) == 1:
            return 0
        if self._symm is not None:
            from .. import ops
            a, b = self.group_ranges[group]
            cur = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)     # this rank's gradients of the group are written
                self._symm.barrier(channel=0)    # ... and everybody else's
                ops.multimem_allreduce_mean_(self._mc_ptr, a, b - a, self._symm.rank, self._symm.world_size, self._mm_ctas)
                self._symm.barrier(channel=0)    # every rank's slice is stored everywhere
                done = torch.cuda.Event()
                done.record(self._side)
            self._works.append((None, done))
            return 1
        avg = dist.get_backend(self.process_group) == "nccl"
        n = 0
        for view in self.buckets(group):
            work = dist.all_reduce(view, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.process_group,
                                   async_op=True)
            self._works.append((work, None if avg else view))
            n += 1
        return n

    def wait(self) -> None:
        """The current stream (CUDA) / the caller (CPU backends) waits for every outstanding collective."""
        world = dist.get_world_size(self.process_group) if dist.is_initialized() else 1
        for work, extra in self._works:
            if work is None:
                torch.cuda.current_stream().wait_event(extra)
                continue
            work.wait()
            if extra is not None:
                extra.div_(world)
        self._works = []
