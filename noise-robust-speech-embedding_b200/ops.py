"""Torch-facing operators of the B200 hot path: thin wrappers that hand raw device pointers and the current
CUDA stream to the C-ABI library (include/nrse_b200.h) and register the calls as ``torch.library`` custom ops
(namespace ``nrse``).  PyTorch is plumbing here -- memory, streams, autograd bookkeeping -- not arithmetic.

Every function requires CUDA tensors on an sm_100 device and raises otherwise: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_F32, NORM_GROUP, NORM_LAYER, FrontendBwdWeights, FrontendGrads, FrontendParams,
                   NrseError, check)

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)

# Parameter generation.  The multi-tensor kernels (EMA, fused AdamW) write parameters through raw device pointers,
# which does NOT bump ``tensor._version``; every such writer bumps this counter instead, and everything derived from
# parameter values (the bf16 weight packs of the conv frontend) is keyed on it.
_param_generation = 0


def param_generation() -> int:
    return _param_generation


def bump_param_generation() -> None:
    global _param_generation
    _param_generation += 1


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> Optional[C.c_void_p]:
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NrseError("nrse_b200 operators need CUDA tensors on a B200 (there is no CPU fallback)")


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise NrseError(f"unsupported dtype {t.dtype}: float32 or bfloat16 expected")


# --------------------------------------------------------------------------------------------------------------
# SNR mix + peak-norm + z-norm
# --------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("nrse::mix_normalize", mutates_args=())
def _mix_normalize_op(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db: Sequence[float],
                      mode: int) -> Tuple[Tensor, Tensor, Tensor]:
    # mode: 1 = BYOL (peak-norm + z-norm, two outputs), 0 = emotion (z-norm only), 2 = raw mix (no normalisation)
    peak_norm = mode == 1
    _need_cuda(clean, noise, snr_idx)
    lib = _lib.load()
    B, L = clean.shape
    noisy_out = torch.empty_like(clean)
    clean_out = torch.empty_like(clean) if peak_norm else clean.new_empty(0)
    status = torch.empty(B, dtype=torch.int32, device=clean.device)
    table = (C.c_double * len(snr_db))(*[float(v) for v in snr_db])
    check(lib.nrse_mix_normalize_f32(_ptr(clean), _ptr(noise), _ptr(snr_idx), table, len(snr_db),
                                     _ptr(clean_out) if peak_norm else None, _ptr(noisy_out), _ptr(status),
                                     B, L, noise.shape[1], int(mode), _stream()),
          "nrse_mix_normalize_f32")
    return clean_out, noisy_out, status


@_mix_normalize_op.register_fake
def _(clean, noise, snr_idx, snr_db, mode):
    return (torch.empty_like(clean) if mode == 1 else clean.new_empty(0), torch.empty_like(clean),
            clean.new_empty(clean.shape[0], dtype=torch.int32))


def mix_normalize(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db_table: Sequence[float],
                  peak_norm: bool = True) -> Tuple[Optional[Tensor], Tensor, Tensor]:
    """Batched ``add_noise_to_speech`` + peak normalisation + HF z-normalisation on the GPU.

    clean [B,L] f32, noise [B,Ln] f32, snr_idx [B] int32 indexing ``snr_db_table`` (dB, e.g. the YAML
    ``data.snr_range``).  Returns (clean_input_values | None, noisy_input_values, status [B] int32);
    ``status[b] != 0`` names the reference's ``return None`` / retry exit that row b would have taken
    (ref:src/data/augment.py:7-64, ref:src/data/noisy_speech_dataset.py:95-138); such rows are zero-filled
    in BYOL mode.  peak_norm=False is the emotion fine-tune variant (ref:src/data/emotion_dataset.py:177-203).
    """
    if clean.dim() != 2 or noise.dim() != 2 or clean.shape[0] != noise.shape[0]:
        raise NrseError("mix_normalize expects clean [B,L] and noise [B,Ln]")
    clean = clean.contiguous().float()
    noise = noise.contiguous().float()
    snr_idx = snr_idx.to(device=clean.device, dtype=torch.int32).contiguous()
    c, n, st = _mix_normalize_op(clean, noise, snr_idx, [float(v) for v in snr_db_table], 1 if peak_norm else 0)
    return (c if peak_norm else None), n, st


@torch.library.custom_op("nrse::mix_normalize_retry", mutates_args=("clean_out", "noisy_out", "status", "snr_used"))
def _mix_retry_op(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db: Sequence[float], mode: int,
                  clean_out: Tensor, noisy_out: Tensor, status: Tensor, snr_used: Tensor, noise_row_shift: int) -> None:
    _need_cuda(clean, noise, snr_idx, noisy_out, status)
    B, L = clean.shape
    table = (C.c_double * len(snr_db))(*[float(v) for v in snr_db])
    check(_lib.load().nrse_mix_normalize_retry_f32(
        _ptr(clean), _ptr(noise), _ptr(snr_idx), table, len(snr_db), _ptr(clean_out) if mode == 1 else None,
        _ptr(noisy_out), _ptr(status), _ptr(snr_used) if snr_used.numel() else None, B, L, noise.shape[1], int(mode),
        int(noise_row_shift), _stream()), "nrse_mix_normalize_retry_f32")


def mix_normalize_retry_(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db_table: Sequence[float],
                         clean_out: Optional[Tensor], noisy_out: Tensor, status: Tensor, noise_row_shift: int,
                         peak_norm: bool = True, snr_idx_used: Optional[Tensor] = None) -> None:
    """In place: redo ``mix_normalize`` for the rows with ``status != 0`` using the noise AND the SNR index of row
    ``(b + noise_row_shift) % B`` -- the reference's next attempt draws another noise file and another SNR
    (ref:src/data/noisy_speech_dataset.py:58-84) -- decided and executed on the device; rows that were fine are not
    touched and cost nothing.  ``snr_idx`` is never modified; ``snr_idx_used`` (int32 [B], seeded by the caller with a
    copy of ``snr_idx``) receives the index every re-done row was finally mixed at."""
    if not (clean.is_contiguous() and noise.is_contiguous() and noisy_out.is_contiguous() and status.is_contiguous()):
        raise NrseError("mix_normalize_retry_ expects contiguous tensors (it works in place)")
    if clean.dtype != torch.float32 or noise.dtype != torch.float32 or status.dtype != torch.int32:
        raise NrseError("mix_normalize_retry_ expects fp32 waveforms and an int32 status")
    if snr_idx_used is not None and (snr_idx_used.dtype != torch.int32 or not snr_idx_used.is_contiguous()
                                     or snr_idx_used.numel() != clean.shape[0]):
        raise NrseError("mix_normalize_retry_: snr_idx_used must be a contiguous int32 [B] tensor")
    snr_idx = snr_idx.to(device=clean.device, dtype=torch.int32).contiguous()
    co = clean_out if peak_norm else clean.new_empty(0)
    su = snr_idx_used if snr_idx_used is not None else status.new_empty(0)
    _mix_retry_op(clean, noise, snr_idx, [float(v) for v in snr_db_table], 1 if peak_norm else 0, co, noisy_out, status,
                  su, int(noise_row_shift))


@torch.library.custom_op("nrse::mix_substitute_rows", mutates_args=("clean_out", "noisy_out", "snr_used"))
def _mix_substitute_op(clean_out: Tensor, noisy_out: Tensor, status: Tensor, snr_used: Tensor) -> None:
    _need_cuda(noisy_out, status)
    B, L = noisy_out.shape
    check(_lib.load().nrse_mix_substitute_rows_f32(
        _ptr(clean_out) if clean_out.numel() else None, _ptr(noisy_out), _ptr(status),
        _ptr(snr_used) if snr_used.numel() else None, B, L, _stream()), "nrse_mix_substitute_rows_f32")


def mix_substitute_rows_(clean_out: Optional[Tensor], noisy_out: Tensor, status: Tensor,
                         snr_idx_used: Optional[Tensor] = None) -> None:
    """In place, no host sync: rows whose ``status`` is still non-zero after the retries take over the outputs (and the
    SNR index) of the nearest following good row -- the reference moves on to the next item instead of emitting an
    unusable one (ref:src/data/noisy_speech_dataset.py:60-66).  On a healthy batch the launch moves no data."""
    if not (noisy_out.is_contiguous() and status.is_contiguous() and (clean_out is None or clean_out.is_contiguous())):
        raise NrseError("mix_substitute_rows_ expects contiguous tensors (it works in place)")
    empty = status.new_empty(0)
    _mix_substitute_op(clean_out if clean_out is not None else noisy_out.new_empty(0), noisy_out, status,
                       snr_idx_used if snr_idx_used is not None else empty)


@torch.library.custom_op("nrse::mix_batch", mutates_args=())
def _mix_batch_op(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db: Sequence[float], label_table: Tensor, mode: int,
                  max_attempts: int, substitute: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _need_cuda(clean, noise, snr_idx, label_table)
    B, L = clean.shape
    noisy_out = torch.empty_like(clean)
    clean_out = torch.empty_like(clean) if mode == 1 else clean.new_empty(0)
    # status [B] | snr index finally used [B] | number of rows rejected for good [1], one int32 allocation
    meta = torch.empty(2 * B + 1, dtype=torch.int32, device=clean.device)
    labels = torch.empty(B, dtype=torch.int64, device=clean.device)
    table = (C.c_double * len(snr_db))(*[float(v) for v in snr_db])
    base = meta.data_ptr()
    check(_lib.load().nrse_mix_batch_f32(
        _ptr(clean), _ptr(noise), _ptr(snr_idx), table, len(snr_db), _ptr(clean_out) if mode == 1 else None,
        _ptr(noisy_out), C.c_void_p(base), C.c_void_p(base + 4 * B), _ptr(label_table), _ptr(labels),
        C.c_void_p(base + 8 * B), B, L, noise.shape[1], int(mode), int(max_attempts), 1 if substitute else 0, _stream()),
        "nrse_mix_batch_f32")
    return clean_out, noisy_out, meta, labels   # (outputs of a custom op may not alias each other: `meta` is sliced outside)


@_mix_batch_op.register_fake
def _(clean, noise, snr_idx, snr_db, label_table, mode, max_attempts, substitute):
    B = clean.shape[0]
    return (torch.empty_like(clean) if mode == 1 else clean.new_empty(0), torch.empty_like(clean),
            clean.new_empty(2 * B + 1, dtype=torch.int32), clean.new_empty(B, dtype=torch.int64))


def mix_batch(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db_table: Sequence[float], snr_label_table: Tensor,
              peak_norm: bool = True, max_attempts: int = 5, substitute: bool = True):
    """The attempt loop of ``NoiseRobustSpeechDataset.__getitem__`` (ref:src/data/noisy_speech_dataset.py:55-149) for a
    whole batch in ONE host call and three launches: mix + normalise every row; redo the rejected rows inside one retry
    launch with the noise and SNR draw of the following rows (up to ``max_attempts - 1`` further attempts each); then
    substitute what is still rejected by the nearest following good row (``substitute``), gather the int64 ``snr`` labels
    (``snr_label_table[index finally used]``, a device int64 table) and count the rows rejected for good.  No host
    synchronisation.  Returns (clean | None, noisy, status [B] i32, snr_idx_used [B] i32, snr_labels [B] i64,
    n_rejected [1] i32)."""
    if clean.dim() != 2 or noise.dim() != 2 or clean.shape[0] != noise.shape[0]:
        raise NrseError("mix_batch expects clean [B,L] and noise [B,Ln]")
    if snr_label_table.dtype != torch.int64 or snr_label_table.numel() != len(snr_db_table):
        raise NrseError("mix_batch: snr_label_table must be an int64 tensor with one entry per SNR table value")
    clean = clean.contiguous().float()
    noise = noise.contiguous().float()
    snr_idx = snr_idx.to(device=clean.device, dtype=torch.int32).contiguous()
    c, n, meta, labels = _mix_batch_op(clean, noise, snr_idx, [float(v) for v in snr_db_table],
                                       snr_label_table.contiguous(), 1 if peak_norm else 0, int(max_attempts), bool(substitute))
    B = clean.shape[0]
    return (c if peak_norm else None), n, meta[:B], meta[B:2 * B], labels, meta[2 * B:]


def mix_status_name(code: int) -> str:
    return _lib.load().nrse_mix_status_name(int(code)).decode()


def mix_raw(clean: Tensor, noise: Tensor, snr_idx: Tensor, snr_db_table: Sequence[float]) -> Tuple[Tensor, Tensor]:
    """``add_noise_to_speech`` alone (ref:src/data/augment.py:4-66), batched: returns (speech + scale*noise [B,L],
    status [B]); rows with status != 0 are the ones for which the reference returns ``None``."""
    clean = clean.contiguous().float()
    noise = noise.contiguous().float()
    snr_idx = snr_idx.to(device=clean.device, dtype=torch.int32).contiguous()
    _, n, st = _mix_normalize_op(clean, noise, snr_idx, [float(v) for v in snr_db_table], 2)
    return n, st


# --------------------------------------------------------------------------------------------------------------
# multi-tensor EMA
# --------------------------------------------------------------------------------------------------------------
class EmaPlan:
    """Chunk table for ``target = decay*target + (1-decay)*online`` over many tensor pairs, built once.

    Replaces the per-tensor loop of ``BYOLSpeechModel._update_target_network`` (ref:src/models/byol.py:62-73).
    The update is in place (the reference re-binds ``.data`` to a fresh tensor every step; values are identical).
    """

    CHUNK_ELEMS = 16384

    def __init__(self, online: Sequence[Tensor], target: Sequence[Tensor]):
        if len(online) != len(target):
            raise NrseError("EmaPlan: online/target lists differ in length")
        self.online = list(online)
        self.target = list(target)
        self._key = None
        self._build()

    def _signature(self):
        return tuple((o.data_ptr(), t.data_ptr(), t.numel()) for o, t in zip(self.online, self.target))

    def _build(self):
        lib = _lib.load()
        n = len(self.online)
        for o, t in zip(self.online, self.target):
            _need_cuda(o, t)
            if o.dtype != torch.float32 or t.dtype != torch.float32:
                raise NrseError("EmaPlan: fp32 parameters expected")
            if o.shape != t.shape or not o.is_contiguous() or not t.is_contiguous():
                raise NrseError("EmaPlan: online/target tensors must be contiguous and of equal shape")
        self._key = self._signature()
        self.n_chunks = 0
        self.numel = sum(t.numel() for t in self.target)
        if n == 0:
            return
        tp = (C.c_uint64 * n)(*[t.data_ptr() for t in self.target])
        op = (C.c_uint64 * n)(*[o.data_ptr() for o in self.online])
        ne = (C.c_int64 * n)(*[t.numel() for t in self.target])
        cnt = lib.nrse_ema_plan_chunks_host(tp, op, ne, n, self.CHUNK_ELEMS, None, None, None, 0)
        if cnt < 0:
            check(int(cnt), "nrse_ema_plan_chunks_host")
        ct = (C.c_uint64 * cnt)()
        co = (C.c_uint64 * cnt)()
        cn = (C.c_int32 * cnt)()
        got = lib.nrse_ema_plan_chunks_host(tp, op, ne, n, self.CHUNK_ELEMS, ct, co, cn, cnt)
        if got != cnt:
            raise NrseError("nrse_ema_plan_chunks_host: inconsistent chunk count")
        dev = self.target[0].device
        # uint64 addresses travel as int64 bit patterns
        self._ct = torch.frombuffer(bytearray(bytes(ct)), dtype=torch.int64).to(dev)
        self._co = torch.frombuffer(bytearray(bytes(co)), dtype=torch.int64).to(dev)
        self._cn = torch.frombuffer(bytearray(bytes(cn)), dtype=torch.int32).to(dev)
        self.n_chunks = int(cnt)

    @torch.no_grad()
    def step(self, decay: float) -> None:
        if self._signature() != self._key:  # parameters were moved / re-allocated
            self._build()
        if self.n_chunks == 0:
            return
        lib = _lib.load()
        # (1 - decay) is evaluated in Python double and then applied as an fp32 scalar, as torch does for
        # `(1 - self.ema_decay) * online_param.data` (ref:src/models/byol.py:67-68)
        check(lib.nrse_ema_chunks_f32(_ptr(self._ct), _ptr(self._co), _ptr(self._cn), self.n_chunks,
                                      float(decay), float(1 - decay), _stream()), "nrse_ema_chunks_f32")
        bump_param_generation()  # targets were written behind autograd's back (no _version bump)


def ema_update_(online: Sequence[Tensor], target: Sequence[Tensor], decay: float) -> None:
    """One-shot convenience wrapper (builds a plan, runs one step)."""
    EmaPlan(online, target).step(decay)


# --------------------------------------------------------------------------------------------------------------
# Fused optimizer tail: clip_grad_norm_ + AdamW + EMA  (ref:train_byol.py:67-71)
# --------------------------------------------------------------------------------------------------------------
class OptimChunkTable:
    """Device chunk table over (param, grad, exp_avg, exp_avg_sq, EMA target twin) for ONE optimizer step count,
    plus the two launches that consume it.  ``update`` rebuilds the table only when an address changed.  A rebuild is
    host planning in C plus ONE non-blocking upload from a pinned staging buffer on the current stream (two staging
    buffers, each guarded by an event), so even a step loop that frees its gradients every iteration
    (``zero_grad(set_to_none=True)``) never synchronises the host with the device."""

    CHUNK_ELEMS = 16384

    def __init__(self):
        self._key = None
        self.n_chunks = 0
        self.numel = 0          # elements that take an AdamW update
        self.ema_numel = 0      # elements that take an EMA update
        self._ptrs = self._numel = None
        self._stage = [None, None]   # (pinned int64 tensor, event)
        self._flip = 0
        self._dev_buf = None

    @staticmethod
    def partials_count() -> int:
        return int(_lib.load().nrse_optim_partials_count())

    def update(self, params: Sequence[Tensor], grads, exp_avgs, exp_avg_sqs, twins) -> None:
        """``grads[i]`` may be None (AdamW skips the tensor; its twin, if any, is still averaged); ``twins[i]`` may
        be None (no EMA for that parameter)."""
        entries, dev = [], None
        for p, g, m, v, t in zip(params, grads, exp_avgs, exp_avg_sqs, twins):
            if g is None and t is None:
                continue
            ts = [x for x in ((p, g, m, v, t) if g is not None else (p, t)) if x is not None]
            _need_cuda(*ts)
            for x in ts:
                if x.dtype != torch.float32 or not x.is_contiguous() or x.numel() != p.numel() or x.is_sparse:
                    raise NrseError("OptimChunkTable: dense contiguous fp32 tensors of the parameter's size expected")
            dev = p.device
            entries.append((p.data_ptr(), 0 if g is None else g.data_ptr(), 0 if g is None else m.data_ptr(),
                            0 if g is None else v.data_ptr(), 0 if t is None else t.data_ptr(), p.numel()))
        key = tuple(entries)
        if key == self._key:
            return
        self._key = key
        lib = _lib.load()
        n = len(entries)
        self.numel = sum(e[5] for e in entries if e[1])
        self.ema_numel = sum(e[5] for e in entries if e[4])
        self.n_chunks = 0
        if n == 0:
            return
        host = torch.tensor(entries, dtype=torch.int64).t().contiguous()  # [6, n]: p | g | m | v | t | numel
        hp = host.data_ptr()
        row = lambda k: C.c_void_p(hp + 8 * n * k)
        cnt = int(lib.nrse_optim_plan_chunks_host(row(0), row(1), row(2), row(3), row(4), row(5), n, self.CHUNK_ELEMS,
                                                  None, None, 0))
        if cnt < 0:
            check(cnt, "nrse_optim_plan_chunks_host")
        if cnt == 0:
            return
        # staging layout (int64 words): 5 * cnt chunk addresses, then cnt int32 lengths packed in (cnt + 1) // 2 words
        words = 5 * cnt + (cnt + 1) // 2
        slot = self._flip
        self._flip ^= 1
        stage = self._stage[slot]
        if stage is None or stage[0].numel() < words:
            stage = (torch.empty(max(words, 1024), dtype=torch.int64).pin_memory(), torch.cuda.Event())
            self._stage[slot] = stage
        else:
            stage[1].synchronize()  # the upload that last read this staging buffer (two rebuilds ago) has finished
        buf, ev = stage
        sp = buf.data_ptr()
        got = lib.nrse_optim_plan_chunks_host(row(0), row(1), row(2), row(3), row(4), row(5), n, self.CHUNK_ELEMS,
                                              C.c_void_p(sp), C.c_void_p(sp + 8 * 5 * cnt), cnt)
        if got != cnt:
            raise NrseError("nrse_optim_plan_chunks_host: inconsistent chunk count")
        if self._dev_buf is None or self._dev_buf.numel() < words or self._dev_buf.device != dev:
            self._dev_buf = torch.empty(max(words, 1024), dtype=torch.int64, device=dev)
        self._dev_buf[:words].copy_(buf[:words], non_blocking=True)  # stream-ordered before the kernels that read it
        ev.record()
        self._ptrs = self._dev_buf[:5 * cnt]
        self._numel = self._dev_buf[5 * cnt:words].view(torch.int32)[:cnt]
        self.n_chunks = cnt

    def grad_sqnorm(self, partials: Tensor) -> None:
        """Writes ``partials_count()`` fp64 partial sums of squares of this table's gradients into ``partials``."""
        g_ptrs = None if self.n_chunks == 0 else C.c_void_p(self._ptrs.data_ptr() + 8 * self.n_chunks)
        check(_lib.load().nrse_grad_sqnorm_chunks_f32(g_ptrs, _ptr(self._numel), self.n_chunks, _ptr(partials),
                                                      _stream()), "nrse_grad_sqnorm_chunks_f32")

    def clip_adamw_ema(self, *, lr: float, betas, eps: float, weight_decay: float, step: int, max_grad_norm: float,
                       ema_decay: float, partials: Optional[Tensor], norm_out: Optional[Tensor]) -> None:
        if self.n_chunks == 0:
            return
        check(_lib.load().nrse_clip_adamw_ema_chunks_f32(
            _ptr(self._ptrs), self.n_chunks, _ptr(self._numel), self.n_chunks, float(lr), float(betas[0]),
            float(betas[1]), float(eps), float(weight_decay), int(step), float(max_grad_norm), float(ema_decay),
            _ptr(partials), 0 if partials is None else partials.numel(), _ptr(norm_out), _stream()),
            "nrse_clip_adamw_ema_chunks_f32")
        bump_param_generation()  # parameters (and EMA twins) were written through raw pointers


# --------------------------------------------------------------------------------------------------------------
# BYOL loss
# --------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("nrse::byol_loss_fwd", mutates_args=())
def _byol_loss_fwd(p: Tensor, z: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _need_cuda(p, z)
    lib = _lib.load()
    B, D = p.shape
    loss = torch.empty((), dtype=torch.float32, device=p.device)
    saved = torch.empty(B, 4, dtype=torch.float32, device=p.device)
    row_sim = torch.empty(B, dtype=torch.float32, device=p.device)
    flags = torch.empty((), dtype=torch.int32, device=p.device)
    check(lib.nrse_byol_loss_fwd(_ptr(p), _ptr(z), _ptr(loss), _ptr(saved), _ptr(row_sim), _ptr(flags), B, D,
                                 _dtype_code(p), _stream()), "nrse_byol_loss_fwd")
    return loss, saved, row_sim, flags


@_byol_loss_fwd.register_fake
def _(p, z):
    B = p.shape[0]
    return (p.new_empty((), dtype=torch.float32), p.new_empty(B, 4, dtype=torch.float32),
            p.new_empty(B, dtype=torch.float32), p.new_empty((), dtype=torch.int32))


@torch.library.custom_op("nrse::byol_loss_bwd", mutates_args=())
def _byol_loss_bwd(p: Tensor, z: Tensor, saved: Tensor, grad_loss: Tensor) -> Tensor:
    lib = _lib.load()
    B, D = p.shape
    grad_p = torch.empty_like(p)
    g = grad_loss.to(torch.float32).contiguous()
    check(lib.nrse_byol_loss_bwd(_ptr(p), _ptr(z), _ptr(saved), _ptr(g), _ptr(grad_p), B, D, _dtype_code(p),
                                 _stream()), "nrse_byol_loss_bwd")
    return grad_p


@_byol_loss_bwd.register_fake
def _(p, z, saved, grad_loss):
    return torch.empty_like(p)


def _loss_setup(ctx, inputs, output):
    p, z = inputs
    ctx.save_for_backward(p, z, output[1])


def _loss_backward(ctx, g_loss, g_saved, g_rowsim, g_flags):
    p, z, saved = ctx.saved_tensors
    return _byol_loss_bwd(p, z, saved, g_loss), None  # no gradient to the target branch (byol.py:94-96)


_byol_loss_fwd.register_autograd(_loss_backward, setup_context=_loss_setup)


def _loss_inputs(online_pred: Tensor, target_proj: Tensor):
    if online_pred.dim() != 2 or online_pred.shape != target_proj.shape:
        raise NrseError("byol_loss expects two [B,D] tensors of equal shape")
    if target_proj.dtype != online_pred.dtype:
        target_proj = target_proj.to(online_pred.dtype)
    return online_pred.contiguous(), target_proj.detach().contiguous()


def byol_loss(online_pred: Tensor, target_proj: Tensor) -> Tensor:
    """Drop-in for ``byol_loss`` (ref:src/models/byol.py:104-129): 0-d fp32 tensor, differentiable w.r.t.
    ``online_pred``; one kernel launch forward, one backward, no host synchronisation."""
    p, z = _loss_inputs(online_pred, target_proj)
    return _byol_loss_fwd(p, z)[0]


def byol_loss_with_flags(online_pred: Tensor, target_proj: Tensor) -> Tuple[Tensor, Tensor]:
    """(loss, flags): ``flags`` is a 0-d int32 device tensor written by the same launch -- bit 0 / 1: NaN in
    ``online_pred`` / ``target_proj`` before normalisation, bit 2 / 3: NaN after it (the two diagnostics of
    ref:src/models/byol.py:109-122).  Reading it is the caller's (only) host synchronisation."""
    p, z = _loss_inputs(online_pred, target_proj)
    out = _byol_loss_fwd(p, z)
    return out[0], out[3]


# --------------------------------------------------------------------------------------------------------------
# Fused tensor health check (ref:src/utils/debugging_utils.py:4-30)
# --------------------------------------------------------------------------------------------------------------
CHECK_MAX_TENSORS = 8
CHECK_NAN, CHECK_INF, CHECK_SMALL, CHECK_LARGE = 1, 2, 4, 8


def check_tensors(tensors: Sequence[Tensor], max_threshold: float = 1e6, min_threshold: float = 1e-6) -> Tensor:
    """One launch over up to 8 tensors.  Returns a device uint8 tensor [n, 64]: one ``nrse_tensor_check`` record per
    tensor (decode with ``decode_tensor_checks`` after copying it to the host -- that copy is the only synchronisation)."""
    if not 1 <= len(tensors) <= CHECK_MAX_TENSORS:
        raise NrseError(f"check_tensors takes 1..{CHECK_MAX_TENSORS} tensors")
    ts = [t.detach().contiguous().float() for t in tensors]
    _need_cuda(*ts)
    n = len(ts)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    numel = (C.c_int64 * n)(*[t.numel() for t in ts])
    out = torch.empty(n, 64, dtype=torch.uint8, device=ts[0].device)
    check(_lib.load().nrse_check_tensors_f32(ptrs, numel, n, float(max_threshold), float(min_threshold), _ptr(out),
                                             _stream()), "nrse_check_tensors_f32")
    return out


def decode_tensor_checks(records_host: Tensor) -> List[dict]:
    """Host copy of ``check_tensors``' output -> list of dicts (flags, abs_max, max, min, abs_sum, mean, std)."""
    raw = bytes(records_host.contiguous().cpu().numpy().tobytes())
    out = []
    for i in range(len(raw) // 64):
        r = _lib.TensorCheck.from_buffer_copy(raw[64 * i:64 * (i + 1)])
        n = max(int(r.numel), 1)
        mean = r.sum / n
        var = (r.sumsq - r.sum * r.sum / n) / (n - 1) if n > 1 else float("nan")  # torch.std: unbiased
        out.append({"flags": int(r.flags), "abs_max": float(r.abs_max), "max": float(r.max), "min": float(r.min),
                    "abs_sum": float(r.abs_sum), "mean": mean, "std": max(var, 0.0) ** 0.5 if var == var else var,
                    "numel": int(r.numel)})
    return out


def cosine_rows(a: Tensor, b: Tensor) -> Tensor:
    """Per-row clamped cosine similarity [B] (what ref:evaluate_byol.py:51-55 computes with F.normalize + sum)."""
    p, z = _loss_inputs(a.detach(), b)
    return _byol_loss_fwd(p, z)[2]


def cosine_rows_plain(a: Tensor, b: Tensor) -> Tensor:
    """Per-row cosine similarity as ref:evaluate_byol.py:51-55 writes it -- ``F.normalize(dim=1)`` on both, row dot
    product, NO clamp: the loss kernel's saved unclamped similarity, same single launch.  Differences to the reference
    expression are confined to degenerate rows: the kernel's +1e-10 offset is below half an ulp of any |x| > 2e-3, and
    its norm floor is 1e-10 where ``F.normalize`` defaults to 1e-12 (only rows whose norm is below 1e-10 can tell)."""
    p, z = _loss_inputs(a.detach().float(), b.float())
    return _byol_loss_fwd(p, z)[1][:, 2]


# --------------------------------------------------------------------------------------------------------------
# Attentive statistics pooling (ref:src/models/pool.py:37-58)
# --------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("nrse::asp_pool_fwd", mutates_args=())
def _asp_pool_fwd(x: Tensor, hl: Tensor, attention: Tensor, lens: Tensor) -> Tuple[Tensor, Tensor]:
    _need_cuda(x, hl, attention, lens)
    B, T, D = x.shape
    out = torch.empty(B, 2 * D, dtype=torch.float32, device=x.device)
    weights = torch.empty(B, T, dtype=torch.float32, device=x.device)
    logits = torch.empty(B, T, dtype=torch.float32, device=x.device)
    check(_lib.load().nrse_asp_pool_fwd(_ptr(x), _ptr(hl), _ptr(attention), _ptr(lens), _ptr(out), _ptr(weights),
                                        _ptr(logits), B, T, D, _stream()), "nrse_asp_pool_fwd")
    return out, weights


@_asp_pool_fwd.register_fake
def _(x, hl, attention, lens):
    B, T, D = x.shape
    return x.new_empty(B, 2 * D, dtype=torch.float32), x.new_empty(B, T, dtype=torch.float32)


@torch.library.custom_op("nrse::asp_pool_bwd", mutates_args=())
def _asp_pool_bwd(x: Tensor, hl: Tensor, attention: Tensor, lens: Tensor, out: Tensor, weights: Tensor,
                  grad_out: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    B, T, D = x.shape
    gx, gh = torch.empty_like(x), torch.empty_like(hl)
    ga = torch.empty(D, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    ws_bytes = int(lib.nrse_asp_pool_bwd_workspace_bytes(B, T, D))
    ws = torch.empty(max(ws_bytes, 4) // 4, dtype=torch.float32, device=x.device)
    g = grad_out.to(torch.float32).contiguous()
    check(lib.nrse_asp_pool_bwd(_ptr(x), _ptr(hl), _ptr(attention), _ptr(lens), _ptr(out), _ptr(weights), _ptr(g),
                                _ptr(gx), _ptr(gh), _ptr(ga), _ptr(ws), ws_bytes, B, T, D, _stream()),
          "nrse_asp_pool_bwd")
    return gx, gh, ga


@_asp_pool_bwd.register_fake
def _(x, hl, attention, lens, out, weights, grad_out):
    return torch.empty_like(x), torch.empty_like(hl), x.new_empty(x.shape[2], dtype=torch.float32)


def _asp_setup(ctx, inputs, output):
    x, hl, attention, lens = inputs
    ctx.save_for_backward(x, hl, attention, lens, output[0], output[1])


def _asp_backward(ctx, g_out, g_weights):
    x, hl, attention, lens, out, weights = ctx.saved_tensors
    gx, gh, ga = _asp_pool_bwd(x, hl, attention, lens, out, weights, g_out)
    return gx, gh, ga, None


_asp_pool_fwd.register_autograd(_asp_backward, setup_context=_asp_setup)


def asp_pool(x: Tensor, hl: Tensor, attention: Tensor, lens: Tensor) -> Tensor:
    """[B,2D] = (weighted mean | weighted std) over the first ``lens[b]`` frames of each utterance, the weights being
    ``softmax_t(tanh(hl) @ attention)`` -- everything of ref:src/models/pool.py:46-57 after ``sap_linear``, for the whole
    batch in two launches (two more backward).  ``x``, ``hl``: [B,T,D] fp32; ``attention``: [D] or [D,1]; ``lens``: [B]."""
    if x.dim() != 3 or x.shape != hl.shape:
        raise NrseError("asp_pool expects x and hl of equal shape [B,T,D]")
    if x.shape[2] % 4 != 0 or x.shape[1] > 4096:
        raise NrseError("asp_pool: D must be a multiple of 4 and T <= 4096")
    att = attention.reshape(-1)
    if att.numel() != x.shape[2]:
        raise NrseError("asp_pool: attention vector length must equal the feature dimension")
    lens = lens.to(device=x.device, dtype=torch.int32).clamp(min=0, max=x.shape[1]).contiguous()
    return _asp_pool_fwd(x.float().contiguous(), hl.float().contiguous(), att.float().contiguous(), lens)[0]


# --------------------------------------------------------------------------------------------------------------
# conv feature encoder
# --------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=256)
def frontend_geometry(n_samples: int) -> Tuple[List[int], List[int]]:
    """(T_i, P_i): valid frames and per-utterance frame pitch of each conv layer's channels-last output."""
    lib = _lib.load()
    T = (C.c_int32 * 7)()
    P = (C.c_int32 * 7)()
    check(lib.nrse_conv_frontend_geometry(int(n_samples), T, P), "nrse_conv_frontend_geometry")
    return list(T), list(P)


def pack_conv_weight(w: Tensor) -> Tensor:
    """[512, 512, k] fp32 checkpoint layout -> bf16 [512, k*512] (K index = tap*512 + c_in) for the tcgen05 kernel."""
    _need_cuda(w)
    lib = _lib.load()
    n, cin, k = w.shape
    if n != 512 or cin != 512:
        raise NrseError("pack_conv_weight expects [512, 512, k]")
    w = w.detach().contiguous().float()
    out = torch.empty(512, k * 512, dtype=torch.bfloat16, device=w.device)
    check(lib.nrse_conv_frontend_pack_weights(_ptr(w), _ptr(out), k, _stream()), "nrse_conv_frontend_pack_weights")
    return out


_workspaces = {}


def _workspace(nbytes: int, device: torch.device) -> Tensor:
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _aligned(ws: Tensor, align: int = 1024) -> Tuple[int, int]:
    base = ws.data_ptr()
    off = (-base) % align
    return base + off, ws.numel() - off


@torch.library.custom_op("nrse::conv_frontend_fwd", mutates_args=())
def _conv_frontend_fwd(x: Tensor, w0: Tensor, w_packed: Sequence[Tensor], gammas: Sequence[Tensor],
                       betas: Sequence[Tensor], norm_mode: int, out_bf16: bool) -> Tensor:
    _need_cuda(x, w0)
    lib = _lib.load()
    B, L = x.shape
    T, P = frontend_geometry(L)
    y = torch.empty(B, P[6], 512, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
    nbytes = lib.nrse_conv_frontend_workspace_bytes(B, L)
    ws = _workspace(nbytes, x.device)
    ws_ptr, ws_len = _aligned(ws)
    prm = FrontendParams()
    prm.w0 = w0.data_ptr()
    for i in range(6):
        prm.w_packed[i] = w_packed[i].data_ptr()
    n_norm = 7 if norm_mode == NORM_LAYER else 1
    for i in range(7):
        prm.gamma[i] = gammas[i].data_ptr() if i < n_norm else None
        prm.beta[i] = betas[i].data_ptr() if i < n_norm else None
    check(lib.nrse_conv_frontend_fwd(_ptr(x), C.byref(prm), norm_mode, _ptr(y), DTYPE_BF16 if out_bf16 else DTYPE_F32,
                                     C.c_void_p(ws_ptr), ws_len, None, B, L, _stream()), "nrse_conv_frontend_fwd")
    return y


def _geometry_py(n_samples: int) -> Tuple[List[int], List[int]]:
    """Pure-Python twin of nrse_conv_frontend_geometry (used for shape inference without touching the library)."""
    T, t = [], int(n_samples)
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        t = (t - k) // s + 1
        T.append(t)
    p6 = max(-(-T[i] // (1 << (6 - i))) for i in range(7))
    return T, [p6 << (6 - i) for i in range(7)]


@_conv_frontend_fwd.register_fake
def _(x, w0, w_packed, gammas, betas, norm_mode, out_bf16):
    _, P = _geometry_py(x.shape[1])
    return x.new_empty(x.shape[0], P[6], 512, dtype=torch.bfloat16 if out_bf16 else torch.float32)


def conv_frontend(x: Tensor, conv_weights: Sequence[Tensor], gammas: Sequence[Optional[Tensor]],
                  betas: Sequence[Optional[Tensor]], norm_mode: str = "layer", out_dtype=torch.float32,
                  packed: Optional[Sequence[Tensor]] = None) -> Tensor:
    """WavLM conv feature encoder forward on the B200 kernels.

    x [B,L] fp32 (z-normalised waveform) -> channels-last features [B, T, 512] (a view of the pitched
    [B, P, 512] kernel output); ``.transpose(1, 2)`` gives the HF layout [B, 512, T] without a copy.
    conv_weights: 7 tensors in checkpoint layout ([512,1,10], then [512,512,k]); gammas/betas: LayerNorm /
    GroupNorm affine parameters per layer (None where the layer has no norm).
    """
    if x.dim() == 3:
        x = x.squeeze(1)
    mode = {"layer": NORM_LAYER, "group": NORM_GROUP}[norm_mode]
    x = x.contiguous().float()
    T, _ = frontend_geometry(x.shape[1])
    w0 = conv_weights[0].detach().reshape(512, 10).contiguous().float()
    if packed is None:
        packed = [pack_conv_weight(w) for w in conv_weights[1:]]
    n_norm = 7 if mode == NORM_LAYER else 1
    g = [gammas[i].detach().contiguous().float() for i in range(n_norm)]
    b = [betas[i].detach().contiguous().float() for i in range(n_norm)]
    y = _conv_frontend_fwd(x, w0, list(packed), g, b, mode, out_dtype == torch.bfloat16)
    return y[:, :T[6], :]


DEFAULT_FRONTEND_VARIANT = 4  # see nrse_conv_frontend_set_variant (include/nrse_b200.h)
DEFAULT_LAYER0_VARIANT = 3    # see nrse_conv_frontend_set_layer0_variant
DEFAULT_BWD_FUSION = False    # see nrse_conv_frontend_set_bwd_fusion


def set_frontend_variant(variant: int) -> None:
    check(_lib.load().nrse_conv_frontend_set_variant(int(variant)), "nrse_conv_frontend_set_variant")


def set_mix_variant(variant: int) -> None:
    check(_lib.load().nrse_mix_set_variant(int(variant)), "nrse_mix_set_variant")


def set_mix_cluster(ctas_per_row: int) -> None:
    check(_lib.load().nrse_mix_set_cluster(int(ctas_per_row)), "nrse_mix_set_cluster")


def set_mix_carveout(percent: int) -> None:
    check(_lib.load().nrse_mix_set_carveout(int(percent)), "nrse_mix_set_carveout")


def multimem_allreduce_mean_(multicast_ptr: int, elem_offset: int, numel: int, rank: int, world: int,
                             max_ctas: int = 0) -> None:
    """In place, on the current stream: mean over the ranks of ``numel`` fp32 elements at ``elem_offset`` of a symmetric
    allocation, through its NVSwitch multicast address (csrc/allreduce.cu).  The caller brackets the call with
    symmetric-memory barriers on the same stream."""
    check(_lib.load().nrse_multimem_allreduce_mean_f32(C.c_void_p(int(multicast_ptr)), int(elem_offset), int(numel), int(rank),
                                                       int(world), int(max_ctas), _stream()),
          "nrse_multimem_allreduce_mean_f32")


def set_sm_budget(sms: int) -> None:
    """SMs the persistent conv-frontend kernels spread over (148 = all).  The data-parallel step lowers it while a
    gradient all-reduce is in flight so that NCCL's kernels find free SMs (see include/nrse_b200.h)."""
    check(_lib.load().nrse_conv_frontend_set_sm_budget(int(sms)), "nrse_conv_frontend_set_sm_budget")


def set_bwd_fusion(on: bool) -> None:
    """LayerNorm / GELU backward of layers 0-5 inside the data-gradient epilogue of the layer above (see the header)."""
    check(_lib.load().nrse_conv_frontend_set_bwd_fusion(1 if on else 0), "nrse_conv_frontend_set_bwd_fusion")


def set_tile_order(alternate: int) -> None:
    check(_lib.load().nrse_conv_frontend_set_tile_order(int(alternate)), "nrse_conv_frontend_set_tile_order")


def set_layer0_variant(variant: int) -> None:
    check(_lib.load().nrse_conv_frontend_set_layer0_variant(int(variant)), "nrse_conv_frontend_set_layer0_variant")


# ---- per-layer entry points (parity tests, callers that own their activation buffers) ---------------------------
def conv_layer0(x: Tensor, w0: Tensor, gamma: Tensor, beta: Tensor, norm_mode: str = "layer") -> Tensor:
    """x [B,L] fp32 -> channels-last bf16 [B, P0, 512] (frames t >= T0 are zero padding)."""
    _need_cuda(x, w0, gamma, beta)
    lib = _lib.load()
    B, L = x.shape
    T, P = frontend_geometry(L)
    out = torch.empty(B, P[0], 512, dtype=torch.bfloat16, device=x.device)
    mode = {"layer": NORM_LAYER, "group": NORM_GROUP}[norm_mode]
    scratch = torch.empty(B * 32 * 1024 * 4 + 2 * (B * 512 * 4 + 1024) + 4096, dtype=torch.uint8, device=x.device)
    sp, _ = _aligned(scratch)
    check(lib.nrse_conv_layer0_fwd(_ptr(x.contiguous()), _ptr(w0.reshape(512, 10).contiguous()), _ptr(gamma), _ptr(beta),
                                   mode, _ptr(out), C.c_void_p(sp), B, L, T[0], P[0], _stream()), "nrse_conv_layer0_fwd")
    return out


def conv_layer(act_prev: Tensor, w_packed: Tensor, k: int, gamma: Optional[Tensor], beta: Optional[Tensor],
               out_dtype=torch.bfloat16) -> Tensor:
    """One stride-2 conv layer as an implicit GEMM: act_prev [rows_prev, 512] bf16 -> [rows_prev/2, 512]."""
    _need_cuda(act_prev, w_packed)
    lib = _lib.load()
    rows_prev = act_prev.shape[0]
    if rows_prev % 2 or act_prev.dtype != torch.bfloat16 or not act_prev.is_contiguous():
        raise NrseError("conv_layer expects a contiguous bf16 [2*rows_out, 512] activation")
    rows_out = rows_prev // 2
    out = torch.empty(rows_out, 512, dtype=out_dtype, device=act_prev.device)
    check(lib.nrse_conv_layer_fwd(_ptr(act_prev), rows_prev, _ptr(w_packed), k, 2, _ptr(gamma), _ptr(beta), _ptr(out),
                                  _dtype_code(out), rows_out, _stream()), "nrse_conv_layer_fwd")
    return out


# ---- training forward / backward of the conv feature encoder (LayerNorm mode) ------------------------------------------
def pack_conv_weight_dgrad(w: Tensor) -> Tuple[Tensor, Tensor]:
    """[512, 512, k] fp32 -> (even [512 c_in, n_even*512], odd [512 c_in, 512]) bf16 operands of the data-gradient GEMMs."""
    _need_cuda(w)
    lib = _lib.load()
    k = w.shape[2]
    w = w.detach().contiguous().float()
    even = torch.empty(512, (2 if k == 3 else 1) * 512, dtype=torch.bfloat16, device=w.device)
    odd = torch.empty(512, 512, dtype=torch.bfloat16, device=w.device)
    check(lib.nrse_conv_frontend_pack_weights_dgrad(_ptr(w), _ptr(even), _ptr(odd), k, _stream()),
          "nrse_conv_frontend_pack_weights_dgrad")
    return even, odd


def _frontend_params(w0: Tensor, packed: Sequence[Tensor], gammas: Sequence[Tensor], betas: Sequence[Tensor]) -> FrontendParams:
    prm = FrontendParams()
    prm.w0 = w0.data_ptr()
    for i in range(6):
        prm.w_packed[i] = packed[i].data_ptr() if packed is not None else None
    for i in range(7):
        prm.gamma[i] = gammas[i].data_ptr() if i < len(gammas) else None
        prm.beta[i] = betas[i].data_ptr() if i < len(betas) else None
    return prm


def _norm_code(norm_mode: str) -> int:
    return {"layer": NORM_LAYER, "group": NORM_GROUP}[norm_mode]


def conv_frontend_train(x: Tensor, conv_weights: Sequence[Tensor], gammas: Sequence[Tensor], betas: Sequence[Tensor],
                        norm_mode: str = "layer", packed: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """Training forward, both norm modes: returns (features [B, T, 512] fp32 view, tape).  The tape (uint8 tensor) holds
    the activations, xhat and the statistics the backward needs; it belongs to the caller (autograd context).
    ``gammas`` / ``betas``: 7 tensors in LayerNorm mode, 1 (layer 0's GroupNorm affine) in GroupNorm mode."""
    _need_cuda(x)
    lib = _lib.load()
    mode = _norm_code(norm_mode)
    n_norm = 7 if mode == NORM_LAYER else 1
    x = x.contiguous().float()
    B, L = x.shape
    T, P = frontend_geometry(L)
    w0 = conv_weights[0].detach().reshape(512, 10).contiguous().float()
    if packed is None:
        packed = [pack_conv_weight(w) for w in conv_weights[1:]]
    g = [gammas[i].detach().contiguous().float() for i in range(n_norm)]
    b = [betas[i].detach().contiguous().float() for i in range(n_norm)]
    nbytes = lib.nrse_conv_frontend_tape_bytes(B, L)
    tape = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
    tp, tl = _aligned(tape)
    y = torch.empty(B, P[6], 512, dtype=torch.float32, device=x.device)
    prm = _frontend_params(w0, packed, g, b)
    check(lib.nrse_conv_frontend_fwd_train(_ptr(x), C.byref(prm), mode, _ptr(y), DTYPE_F32, C.c_void_p(tp), tl, B, L,
                                           _stream()), "nrse_conv_frontend_fwd_train")
    return y[:, :T[6], :], tape


def conv_frontend_backward(x: Tensor, conv_weights: Sequence[Tensor], gammas: Sequence[Tensor], betas: Sequence[Tensor],
                           tape: Tensor, grad_features: Tensor, norm_mode: str = "layer",
                           dgrad_packs: Optional[Sequence[Tuple[Tensor, Tensor]]] = None,
                           need_w: Optional[Sequence[bool]] = None, need_affine: Optional[Sequence[bool]] = None):
    """Backward of ``conv_frontend_train``.  grad_features [B, T, 512] (any strides).  Returns
    (dw: 7 tensors in checkpoint layout [512,1,10] / [512,512,k], dgamma: n_norm x [512], dbeta: n_norm x [512]), all
    fp32 views of ONE zero-initialised buffer; entries that were not asked for (``need_w[i]`` / ``need_affine[i]``
    False: frozen parameters, ref:src/models/emotion.py:114-129) are ``None`` and cost nothing -- the kernels stop at
    the lowest layer that wants a gradient and skip the weight-gradient GEMM of every layer that does not."""
    lib = _lib.load()
    mode = _norm_code(norm_mode)
    n_norm = 7 if mode == NORM_LAYER else 1
    need_w = [True] * 7 if need_w is None else [bool(v) for v in need_w]
    need_affine = [True] * n_norm if need_affine is None else [bool(v) for v in need_affine]
    x = x.contiguous().float()
    B, L = x.shape
    T, P = frontend_geometry(L)
    dev = x.device
    if tape is None:
        raise NrseError("conv_frontend_backward: the tape of this forward has already been consumed")
    if tuple(grad_features.shape) != (B, T[6], 512):
        raise NrseError(f"conv_frontend_backward: grad_features must be [{B}, {T[6]}, 512]")
    dy = grad_features.contiguous().float()
    if not any(need_w) and not any(need_affine):
        return [None] * 7, [None] * n_norm, [None] * n_norm
    lowest = min([i for i, v in enumerate(need_w) if v] + [i for i, v in enumerate(need_affine) if v])
    w0 = conv_weights[0].detach().reshape(512, 10).contiguous().float()
    g = [gammas[i].detach().contiguous().float() for i in range(n_norm)]
    b = [betas[i].detach().contiguous().float() for i in range(n_norm)]
    if dgrad_packs is None:  # only the layers above the lowest one propagate a data gradient
        dgrad_packs = [pack_conv_weight_dgrad(conv_weights[i]) if i > lowest else None for i in range(1, 7)]
    prm = _frontend_params(w0, None, g, b)
    wb = FrontendBwdWeights()
    for i in range(6):
        pk = dgrad_packs[i]
        wb.wt_even[i] = pk[0].data_ptr() if pk is not None else None
        wb.wt_odd[i] = pk[1].data_ptr() if pk is not None else None
    sizes_w = [512 * 10] + [512 * 512 * CONV_KERNEL[i] for i in range(1, 7)]
    total = sum(sz for sz, nd in zip(sizes_w, need_w) if nd) + 1024 * sum(need_affine)
    flat = torch.zeros(total, dtype=torch.float32, device=dev)  # the kernels accumulate: one memset for everything
    off = 0
    dw: List[Optional[Tensor]] = []
    for i in range(7):
        if need_w[i]:
            shape = (512, 1, 10) if i == 0 else (512, 512, CONV_KERNEL[i])
            dw.append(flat[off:off + sizes_w[i]].view(shape))
            off += sizes_w[i]
        else:
            dw.append(None)
    dgam: List[Optional[Tensor]] = []
    dbet: List[Optional[Tensor]] = []
    for i in range(n_norm):
        if need_affine[i]:
            dgam.append(flat[off:off + 512])
            dbet.append(flat[off + 512:off + 1024])
            off += 1024
        else:
            dgam.append(None)
            dbet.append(None)
    gr = FrontendGrads()
    gr.dw0 = dw[0].data_ptr() if dw[0] is not None else None
    for i in range(6):
        gr.dw[i] = dw[i + 1].data_ptr() if dw[i + 1] is not None else None
    for i in range(7):
        gr.dgamma[i] = dgam[i].data_ptr() if i < n_norm and dgam[i] is not None else None
        gr.dbeta[i] = dbet[i].data_ptr() if i < n_norm and dbet[i] is not None else None
    nbytes = lib.nrse_conv_frontend_bwd_workspace_bytes(B, L)
    ws = _workspace(nbytes, dev)
    wp, wl = _aligned(ws)
    tp, _ = _aligned(tape)
    check(lib.nrse_conv_frontend_bwd(_ptr(x), C.byref(prm), C.byref(wb), mode, C.c_void_p(tp), _ptr(dy), T[6],
                                     C.byref(gr), C.c_void_p(wp), wl, B, L, _stream()), "nrse_conv_frontend_bwd")
    return dw, dgam, dbet


# per-kernel hooks (parity tests)
def ln_gelu_bwd(dout: Tensor, xhat: Tensor, rstd: Optional[Tensor], gamma: Optional[Tensor], beta: Optional[Tensor],
                P: int, T: int, dout_pitch: Optional[int] = None):
    """dOut -> dZ.  ``gamma is None``: the no-norm form (GroupNorm-mode layers 1-6), ``xhat`` = the pre-GELU activation."""
    lib = _lib.load()
    rows = xhat.shape[0]
    dz = torch.empty(rows, 512, dtype=torch.bfloat16, device=xhat.device)
    dg = torch.zeros(512, dtype=torch.float32, device=xhat.device) if gamma is not None else None
    db = torch.zeros(512, dtype=torch.float32, device=xhat.device) if gamma is not None else None
    check(lib.nrse_ln_gelu_bwd(_ptr(dout.contiguous()), _dtype_code(dout), int(dout_pitch if dout_pitch is not None else P),
                               _ptr(xhat), _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(dz), _ptr(dg), _ptr(db), rows, P, T,
                               _stream()), "nrse_ln_gelu_bwd")
    return dz, dg, db


def conv_layer_wgrad(dz: Tensor, act_prev: Tensor, k: int, ckpt_layout: bool = False) -> Tensor:
    """dW = dZ^T A: fp32 [512, k*512] in the packed K order, or [512, 512, k] (checkpoint layout)."""
    lib = _lib.load()
    shape = (512, 512, k) if ckpt_layout else (512, k * 512)
    dw = torch.zeros(shape, dtype=torch.float32, device=dz.device)
    check(lib.nrse_conv_layer_wgrad(_ptr(dz), _ptr(act_prev), dz.shape[0], k, _ptr(dw), 1 if ckpt_layout else 0,
                                    _stream()), "nrse_conv_layer_wgrad")
    return dw


def conv_layer_dgrad(dz: Tensor, wt_even: Tensor, wt_odd: Tensor, k: int) -> Tensor:
    lib = _lib.load()
    dx = torch.empty(2 * dz.shape[0], 512, dtype=torch.bfloat16, device=dz.device)
    check(lib.nrse_conv_layer_dgrad(_ptr(dz), dz.shape[0], _ptr(wt_even), _ptr(wt_odd), k, _ptr(dx), _stream()),
          "nrse_conv_layer_dgrad")
    return dx


def conv_layer_dgrad_lnbwd(dz: Tensor, wt_even: Tensor, wt_odd: Tensor, k: int, xhat_prev: Tensor, rstd_prev: Tensor,
                           gamma_prev: Tensor, beta_prev: Tensor, P_prev: int, T_prev: int, want_affine: bool = True):
    """Data gradient of layer i fused with the LayerNorm + GELU backward of layer i-1 -> (dZ_{i-1}, dgamma, dbeta)."""
    lib = _lib.load()
    dzp = torch.empty(2 * dz.shape[0], 512, dtype=torch.bfloat16, device=dz.device)
    dg = torch.zeros(512, dtype=torch.float32, device=dz.device) if want_affine else None
    db = torch.zeros(512, dtype=torch.float32, device=dz.device) if want_affine else None
    check(lib.nrse_conv_layer_dgrad_lnbwd(_ptr(dz), dz.shape[0], _ptr(wt_even), _ptr(wt_odd), k, _ptr(xhat_prev),
                                          _ptr(rstd_prev), _ptr(gamma_prev), _ptr(beta_prev), _ptr(dzp), _ptr(dg),
                                          _ptr(db), int(P_prev), int(T_prev), _stream()), "nrse_conv_layer_dgrad_lnbwd")
    return dzp, dg, db


def conv_layer0_wgrad(x: Tensor, dz0: Tensor, T0: int, P0: int) -> Tensor:
    lib = _lib.load()
    B, L = x.shape
    dw0 = torch.zeros(512, 10, dtype=torch.float32, device=x.device)
    check(lib.nrse_conv_layer0_wgrad(_ptr(x), _ptr(dz0), _ptr(dw0), B, L, T0, P0, _stream()), "nrse_conv_layer0_wgrad")
    return dw0


# --------------------------------------------------------------------------------------------------------------
# Feature projection: LayerNorm(512) + Linear(512 -> 1024)  (hf:models/wavlm/modeling_wavlm.py:93-105)
# --------------------------------------------------------------------------------------------------------------
def pack_feature_projection(weight: Tensor) -> Tuple[Tensor, Tensor]:
    """projection.weight [1024, 512] fp32 -> (bf16 copy, bf16 transpose [512, 1024]): forward / backward GEMM operands."""
    _need_cuda(weight)
    if tuple(weight.shape) != (1024, 512):
        raise NrseError("pack_feature_projection expects projection.weight of shape [1024, 512]")
    w = weight.detach().contiguous().float()
    w16 = torch.empty(1024, 512, dtype=torch.bfloat16, device=w.device)
    wt16 = torch.empty(512, 1024, dtype=torch.bfloat16, device=w.device)
    check(_lib.load().nrse_feature_projection_pack(_ptr(w), _ptr(w16), _ptr(wt16), _stream()), "nrse_feature_projection_pack")
    return w16, wt16


def _channels_last_rows(x: Tensor) -> Tuple[Tensor, int]:
    """[B, T, 512] -> (tensor whose frames (b, t) sit at row b * pitch + t of a dense [.., 512] buffer, pitch): the conv
    frontend hands out a [B, T, 512] view of its pitched [B, P, 512] output, which is read in place."""
    B, T, Cc = x.shape
    if x.stride(2) == 1 and x.stride(1) == Cc and x.stride(0) % Cc == 0 and x.stride(0) >= T * Cc and \
            x.data_ptr() % 16 == 0:
        return x, x.stride(0) // Cc
    return x.contiguous(), T


def feature_projection_fwd(feats: Tensor, ln_weight: Tensor, ln_bias: Tensor, eps: float, w16: Tensor, bias: Tensor,
                           training: bool, want_norm: bool = True) -> Tuple[Tensor, Optional[Tensor], Tensor]:
    """feats [B, T, 512] (fp32 / bf16, channels-last; a pitched view is read in place) -> (hidden [B, T, 1024] fp32,
    norm_hidden [B, T, 512] fp32 | None, tape)."""
    lib = _lib.load()
    _need_cuda(feats, w16)
    if feats.dim() != 3 or feats.shape[2] != 512:
        raise NrseError("feature_projection_fwd expects [B, T, 512] channels-last conv features")
    if feats.dtype not in (torch.float32, torch.bfloat16):
        feats = feats.float()
    x, pitch = _channels_last_rows(feats)
    B, T, _ = x.shape
    rows = B * T
    hidden = torch.empty(B, T, 1024, dtype=torch.float32, device=x.device)
    norm = torch.empty(B, T, 512, dtype=torch.float32, device=x.device) if want_norm else None
    tape = torch.empty(lib.nrse_feature_projection_tape_bytes(rows) + 1024, dtype=torch.uint8, device=x.device)
    tp, _ = _aligned(tape)
    check(lib.nrse_feature_projection_fwd(_ptr(x), _dtype_code(x), B, T, pitch, _ptr(ln_weight.detach().contiguous().float()),
                                          _ptr(ln_bias.detach().contiguous().float()), float(eps), _ptr(w16),
                                          _ptr(bias.detach().contiguous().float()), _ptr(hidden), _ptr(norm),
                                          C.c_void_p(tp), 1 if training else 0, _stream()), "nrse_feature_projection_fwd")
    return hidden, norm, tape


def feature_projection_bwd(d_hidden: Tensor, tape: Tensor, ln_weight: Tensor, ln_bias: Tensor, wt16: Tensor,
                           need_feats: bool = True, need_ln: bool = True, need_w: bool = True, need_bias: bool = True):
    """Backward of ``feature_projection_fwd(training=True)``: d_hidden [B, T, 1024] -> (d_feats [B, T, 512] fp32 | None,
    d_ln_weight, d_ln_bias, d_weight [1024, 512], d_bias [1024]); entries not asked for are None and cost nothing."""
    lib = _lib.load()
    B, T, _ = d_hidden.shape
    rows = B * T
    dev = d_hidden.device
    dh = d_hidden.contiguous().float()
    d_feats = torch.empty(B, T, 512, dtype=torch.float32, device=dev) if need_feats else None
    n_acc = (1024 if need_ln else 0) + (1024 * 512 if need_w else 0) + (1024 if need_bias else 0)
    flat = torch.zeros(max(n_acc, 4), dtype=torch.float32, device=dev)  # the kernels accumulate
    off = 0
    d_g = d_b = d_w = d_bias = None
    if need_ln:
        d_g, d_b = flat[0:512], flat[512:1024]
        off = 1024
    if need_w:
        d_w = flat[off:off + 1024 * 512].view(1024, 512)
        off += 1024 * 512
    if need_bias:
        d_bias = flat[off:off + 1024]
    ws = _workspace(lib.nrse_feature_projection_bwd_workspace_bytes(rows), dev)
    wp, _ = _aligned(ws)
    tp, _ = _aligned(tape)
    check(lib.nrse_feature_projection_bwd(_ptr(dh), C.c_void_p(tp), _ptr(ln_weight.detach().contiguous().float()),
                                          _ptr(ln_bias.detach().contiguous().float()), _ptr(wt16), _ptr(d_feats), _ptr(d_g),
                                          _ptr(d_b), _ptr(d_w), _ptr(d_bias), C.c_void_p(wp), rows, _stream()),
          "nrse_feature_projection_bwd")
    return d_feats, d_g, d_b, d_w, d_bias


# --------------------------------------------------------------------------------------------------------------
# Positional convolution embedding (hf:models/wavlm/modeling_wavlm.py:48-90)
# --------------------------------------------------------------------------------------------------------------
def pack_pos_conv(v: Tensor, g: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """weight-norm parameters (v [1024, 64, 128], g [1, 1, 128]) -> (forward pack, data-gradient pack, ||v||^2 per tap)."""
    _need_cuda(v, g)
    if tuple(v.shape) != (1024, 64, 128) or g.numel() != 128:
        raise NrseError("pack_pos_conv expects v [1024, 64, 128] and g with 128 elements (wavlm-large positional conv)")
    lib = _lib.load()
    n = lib.nrse_pos_conv_pack_bytes()
    wf = torch.empty(n // 2, dtype=torch.bfloat16, device=v.device)
    wb = torch.empty(n // 2, dtype=torch.bfloat16, device=v.device)
    normsq = torch.empty(128, dtype=torch.float32, device=v.device)
    check(lib.nrse_pos_conv_pack(_ptr(v.detach().contiguous().float()), _ptr(g.detach().reshape(-1).contiguous().float()),
                                 _ptr(wf), _ptr(wb), _ptr(normsq), _stream()), "nrse_pos_conv_pack")
    return wf, wb, normsq


def pos_conv_fwd(x: Tensor, w_fwd: Tensor, bias: Tensor, training: bool) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """x [B, T, 1024] -> (y [B, T, 1024] fp32 = GELU(grouped conv), bf16 copy of x, pre-GELU activation bf16 | None)."""
    _need_cuda(x, w_fwd)
    if x.dim() != 3 or x.shape[2] != 1024:
        raise NrseError("pos_conv_fwd expects hidden states [B, T, 1024]")
    x = x.contiguous().float()
    B, T, _ = x.shape
    y = torch.empty_like(x)
    xb = torch.empty(B * T, 1024, dtype=torch.bfloat16, device=x.device)
    z = torch.empty(B * T, 1024, dtype=torch.bfloat16, device=x.device) if training else None
    check(_lib.load().nrse_pos_conv_fwd(_ptr(x), _ptr(w_fwd), _ptr(bias.detach().contiguous().float()), _ptr(y), _ptr(xb),
                                        _ptr(z), B, T, _stream()), "nrse_pos_conv_fwd")
    return y, xb, z


def pos_conv_bwd(d_y: Tensor, xb: Tensor, z: Tensor, w_bwd: Tensor, v: Tensor, g: Tensor, normsq: Tensor,
                 need_x: bool = True, need_w: bool = True, need_bias: bool = True):
    """-> (d_x [B, T, 1024] | None, d_v [1024, 64, 128] | None, d_g [1, 1, 128] | None, d_bias [1024] | None)."""
    lib = _lib.load()
    B, T, _ = d_y.shape
    dev = d_y.device
    dy = d_y.contiguous().float()
    d_x = torch.empty(B, T, 1024, dtype=torch.float32, device=dev) if need_x else None
    flat = torch.zeros((1024 * 64 * 128 + 128 if need_w else 0) + (1024 if need_bias else 0) + 4, dtype=torch.float32,
                       device=dev)  # the kernels accumulate
    off = 0
    d_v = d_g = d_b = None
    if need_w:
        d_v = flat[:1024 * 64 * 128].view(1024, 64, 128)
        d_g = flat[1024 * 64 * 128:1024 * 64 * 128 + 128].view(1, 1, 128)
        off = 1024 * 64 * 128 + 128
    if need_bias:
        d_b = flat[off:off + 1024]
    ws = _workspace(lib.nrse_pos_conv_bwd_workspace_bytes(B, T), dev)
    wp, _ = _aligned(ws)
    check(lib.nrse_pos_conv_bwd(_ptr(dy), _ptr(xb), _ptr(z), _ptr(w_bwd), _ptr(v.detach().contiguous().float()),
                                _ptr(g.detach().reshape(-1).contiguous().float()), _ptr(normsq), _ptr(d_x), _ptr(d_v),
                                _ptr(d_g), _ptr(d_b), C.c_void_p(wp), B, T, _stream()), "nrse_pos_conv_bwd")
    return d_x, d_v, d_g, d_b
