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

The kernels work from a device table of raw addresses, rebuilt whenever a parameter or gradient address changes
(``table_builds`` counts).  ``zero_grad(set_to_none=True)`` frees the gradients every step and the allocator does not
always hand the same blocks back, so a rebuild is made cheap instead of rare: host planning in C and one non-blocking
upload from pinned memory on the current stream -- no host/device synchronisation.  ``keep_grads=True`` makes
``zero_grad()`` zero in place instead (stable addresses, but autograd then accumulates into the zeros: one extra
read-modify-write per parameter in the backward).
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
        self._tables: Dict[int, ops.OptimChunkTable] = {}
        self._buckets: List[dict] = []
        self._gkey = None
        self.keep_grads = False  # True: zero_grad() zeroes in place, addresses (and tables) never change
        self.table_builds = 0  # how often the address tables were (re)built; steady state: stays constant
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
        self._gkey = None

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

    def zero_grad(self, set_to_none: Optional[bool] = None) -> None:
        """``set_to_none`` defaults to ``not self.keep_grads`` (torch's default, True, unless ``keep_grads`` is set)."""
        super().zero_grad(set_to_none=(not self.keep_grads) if set_to_none is None else set_to_none)

    # ---- step counters -------------------------------------------------------------------------------------------
    # torch keeps one CPU tensor ``state['step']`` per parameter and increments each of them every step (~500 host ops
    # for WavLM-large).  Here the counts live as Python ints per bucket and are written into ``state['step']`` only when
    # somebody looks: ``state_dict()`` / ``sync_state()``.
    def sync_state(self) -> None:
        """Write the current step counts into ``state[p]['step']`` (what torch.optim.AdamW would hold)."""
        for bucket in self._buckets:
            for p in bucket["params"]:
                self.state[p]["step"] = torch.tensor(float(bucket["step"]), dtype=torch.float32)

    def state_dict(self):
        self.sync_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        self._buckets = []  # the loaded step counts win: nothing to flush
        super().load_state_dict(state_dict)
        self._gkey = None   # rebuild the buckets from the loaded step counts

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._gkey = None

    def _rebuild(self, params, grads):
        """Bucket the parameters by (group, step count): one table / kernel call per bucket."""
        self.table_builds += 1
        self.sync_state()
        buckets = {}
        orphans = []  # no gradient this step, but an EMA twin
        group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        for p, g in zip(params, grads):
            twin = self._twins.get(id(p))
            if g is None:
                if twin is not None:
                    orphans.append((p.data, None, None, None, twin))
                continue
            state = self.state[p]
            if len(state) == 0:  # torch.optim.Adam._init_group
                state["step"] = torch.tensor(0.0, dtype=torch.float32)
                state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            key = (group_of[id(p)], int(state["step"]))
            b = buckets.setdefault(key, {"group": key[0], "step": key[1], "params": [], "rows": []})
            b["params"].append(p)
            b["rows"].append((p.data, g, state["exp_avg"], state["exp_avg_sq"], twin))
        self._buckets = [buckets[k] for k in sorted(buckets)]
        if orphans:  # ride along with the first bucket (or alone, with a dummy step count)
            if not self._buckets:
                self._buckets = [{"group": 0, "step": 0, "params": [], "rows": []}]
            self._buckets[0]["rows"] = self._buckets[0]["rows"] + orphans
        for slot, b in enumerate(self._buckets):
            tab = self._tables.get(slot)
            if tab is None:
                tab = self._tables[slot] = ops.OptimChunkTable()
            tab.update(*zip(*b["rows"]))
            b["table"] = tab
            del b["rows"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize"):
                raise RuntimeError("FusedAdamWEma: amsgrad / maximize are not supported")
        params = [p for g in self.param_groups for p in g["params"]]
        grads = [p.grad for p in params]
        # the chunk tables hold raw addresses: rebuild when any parameter / gradient address (or the set of parameters
        # that have a gradient) changed.  Moments are owned by this optimizer and only change via load_state_dict.
        gkey = tuple([p.data_ptr() for p in params] + [0 if g is None else g.data_ptr() for g in grads])
        if gkey != self._gkey:
            self._rebuild(params, grads)
            self._gkey = gkey
        if not self._buckets:
            return loss
        dev = params[0].device
        clip = self.max_grad_norm > 0
        count = ops.OptimChunkTable.partials_count()
        n_b = len(self._buckets)
        if clip and (self._partials is None or self._partials.numel() != count * n_b or self._partials.device != dev):
            self._partials = torch.zeros(count * n_b, dtype=torch.float64, device=dev)
            self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        if clip:
            for slot, b in enumerate(self._buckets):
                b["table"].grad_sqnorm(self._partials[slot * count:(slot + 1) * count])
        for b in self._buckets:
            if b["params"]:
                b["step"] += 1
            group = self.param_groups[b["group"]]
            b["table"].clip_adamw_ema(lr=group["lr"], betas=group["betas"], eps=group["eps"],
                                      weight_decay=group["weight_decay"], step=max(b["step"], 1),
                                      max_grad_norm=self.max_grad_norm,
                                      ema_decay=self.ema_decay if self.ema_decay is not None else 0.0,
                                      partials=self._partials if clip else None, norm_out=self._norm if clip else None)
        self.last_grad_norm = self._norm[0] if clip else None
        return loss
