"""Double-buffered host->device staging on a dedicated copy stream, so that the H2D copy of batch i+1 overlaps the
kernels of batch i (the reference copies with blocking ``.to(device)`` on the compute stream, ref:train_byol.py:49-50).
Used by ``bench.py``'s end-to-end arm and usable around any loader that yields dicts of pinned CPU tensors."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional

import torch


class DevicePrefetcher:
    def __init__(self, device, depth: int = 2):
        if depth < 2:
            raise ValueError("DevicePrefetcher needs depth >= 2 (one slot being consumed, one being filled)")
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._free = [torch.cuda.Event() for _ in range(depth)]
        self._n_put = 0
        self._n_get = 0

    def put(self, host_batch: Dict[str, torch.Tensor]) -> None:
        """Enqueue the H2D copy of one batch (pinned tensors) on the copy stream."""
        i = self._n_put % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self._n_put >= self.depth:
                self.copy_stream.wait_event(self._free[i])  # the slot's previous batch has been consumed
            slot = self._slots[i]
            if slot is None or any(slot[k].shape != v.shape for k, v in host_batch.items()):
                slot = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_batch.items()}
                self._slots[i] = slot
            for k, v in host_batch.items():
                slot[k].copy_(v, non_blocking=True)
            self._ready[i].record(self.copy_stream)
        self._n_put += 1

    def get(self) -> Dict[str, torch.Tensor]:
        """Device tensors of the oldest enqueued batch; the current stream waits for its copy.  Call ``release()``
        after the kernels that read it have been enqueued."""
        assert self._n_get < self._n_put, "get() without a matching put()"
        i = self._n_get % self.depth
        torch.cuda.current_stream(self.device).wait_event(self._ready[i])
        return self._slots[i]

    @property
    def current_slot(self) -> int:
        """Index of the slot the next ``get()`` returns (slot tensors keep their addresses: CUDA-graph friendly)."""
        return self._n_get % self.depth

    def release(self) -> None:
        i = self._n_get % self.depth
        self._free[i].record(torch.cuda.current_stream(self.device))
        self._n_get += 1

    def iterate(self, host_batches: Iterable[Dict[str, torch.Tensor]]) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(host_batches)
        pending = 0
        for _ in range(self.depth - 1):
            try:
                self.put(next(it)); pending += 1
            except StopIteration:
                break
        while pending:
            try:
                self.put(next(it)); pending += 1
            except StopIteration:
                pass
            batch = self.get()
            yield batch
            self.release()
            pending -= 1
