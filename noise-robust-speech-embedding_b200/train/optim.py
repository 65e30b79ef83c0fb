"""``FusedAdamWEma``: the optimizer tail of the BYOL step as two kernel launches.

The reference's step body does, in this order (ref:train_byol.py:67-71):
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    optimizer.step()                      # torch.optim.AdamW(model.parameters(), lr, weight_decay)  ref:train_byol.py:146
    model._update_target_network()        # ref:src/models/byol.py:62-73
``FusedAdamWEma`` IS a ``torch.optim.AdamW`` (same constructor arguments, ``param_groups``, per-parameter ``state``
with ``step`` / ``exp_avg`` / ``exp_avg_sq``, so ``state_dict()`` / ``load_state_dict()`` and LR schedulers are
interchangeable with the reference's checkpoints, ref:train_byol.py:188-196) whose ``step()`` runs
``ops.OptimChunkTable``: one read of every gradient for the global norm, then one pass that clips, applies AdamW and
averages the updated parameter into its EMA target twin.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .. import ops


class FusedAdamWEma(torch.optim.AdamW):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, *, max_grad_norm: float = 0.0,
                 ema_pairs: Optional[Iterable[Tuple[torch.Tensor, torch.Tensor]]] = None,
                 ema_decay: Optional[float] = None):
        """``max_grad_norm`` > 0 folds ``clip_grad_norm_(all params of this optimizer, max_grad_norm)`` into the step.
        ``ema_pairs`` = iterable of (online parameter, target tensor): after its AdamW update each online parameter is
        averaged into its target with ``ema_decay`` (``BYOLSpeechModel._update_target_network``); online parameters
        that are not owned by this optimizer are rejected."""
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=False, fused=False)
        self.max_grad_norm = float(max_grad_norm)
        self.ema_decay = ema_decay
        self._twins: Dict[int, torch.Tensor] = {}
        self._tables: Dict[Tuple[int, int], ops.OptimChunkTable] = {}
        self._partials = None
        self._norm = None
        self.last_grad_norm: Optional[torch.Tensor] = None  # device tensor, what clip_grad_norm_ would have returned
        if ema_pairs is not None:
            self.attach_ema(ema_pairs, ema_decay)

    def attach_ema(self, pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]], decay: float) -> None:
        if decay is None:
            raise ValueError("FusedAdamWEma: ema_decay is required with ema_pairs")
        owned = {id(p) for g in self.param_groups for p in g["params"]}
        self._twins = {}
        for online, target in pairs:
            if id(online) not in owned:
                raise ValueError("FusedAdamWEma: an EMA source parameter is not owned by this optimizer")
            if online.shape != target.shape:
                raise ValueError("FusedAdamWEma: EMA pair with different shapes")
            self._twins[id(online)] = target
        self.ema_decay = float(decay)

    @property
    def has_ema(self) -> bool:
        return bool(self._twins)

    @classmethod
    def for_byol(cls, model, lr: float, weight_decay: float, max_grad_norm: float = 1.0, **kw) -> "FusedAdamWEma":
        """The reference's ``optim.AdamW(model.parameters(), lr=..., weight_decay=...)`` (ref:train_byol.py:146) with
        the clip (ref:train_byol.py:67) and the encoder + projector EMA (ref:src/models/byol.py:62-73) folded in."""
        inner = model.module if hasattr(model, "module") else model
        online, target = inner._ema_pairs()
        return cls(model.parameters(), lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm,
                   ema_pairs=zip(online, [t.data for t in target]), ema_decay=inner.ema_decay, **kw)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # bucket the parameters by (group, step count): one kernel call per bucket, the norm partials side by side
        buckets: Dict[Tuple[int, int], List[tuple]] = {}
        orphans: List[tuple] = []  # no gradient this step, but an EMA twin
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise RuntimeError("FusedAdamWEma: amsgrad / maximize are not supported")
            for p in group["params"]:
                twin = self._twins.get(id(p))
                if p.grad is None:
                    if twin is not None:
                        orphans.append((p.data, None, None, None, twin))
                    continue
                state = self.state[p]
                if len(state) == 0:  # torch.optim.Adam._init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                buckets.setdefault((gi, int(state["step"])), []).append(
                    (p.data, p.grad, state["exp_avg"], state["exp_avg_sq"], twin))
        if not buckets and not orphans:
            return loss
        keys = sorted(buckets)
        if orphans:  # ride along with the first bucket (or alone, with a dummy step count)
            if keys:
                buckets[keys[0]] = buckets[keys[0]] + orphans
            else:
                keys = [(-1, 1)]
                buckets[keys[0]] = orphans
        dev = buckets[keys[0]][0][0].device
        clip = self.max_grad_norm > 0
        count = ops.OptimChunkTable.partials_count()
        if clip and (self._partials is None or self._partials.numel() != count * len(keys) or self._partials.device != dev):
            self._partials = torch.zeros(count * len(keys), dtype=torch.float64, device=dev)
            self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        tables = []
        for slot, key in enumerate(keys):
            tab = self._tables.get((key[0], slot))
            if tab is None:
                tab = self._tables[(key[0], slot)] = ops.OptimChunkTable()
            tab.update(*zip(*buckets[key]))
            tables.append(tab)
            if clip:
                tab.grad_sqnorm(self._partials[slot * count:(slot + 1) * count])
        for key, tab in zip(keys, tables):
            group = self.param_groups[max(key[0], 0)]
            tab.clip_adamw_ema(lr=group["lr"], betas=group["betas"], eps=group["eps"],
                               weight_decay=group["weight_decay"], step=key[1], max_grad_norm=self.max_grad_norm,
                               ema_decay=self.ema_decay if self.ema_decay is not None else 0.0,
                               partials=self._partials if clip else None, norm_out=self._norm if clip else None)
        self.last_grad_norm = self._norm[0] if clip else None
        return loss
