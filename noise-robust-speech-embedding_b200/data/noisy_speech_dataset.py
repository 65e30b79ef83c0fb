"""Data side of the BYOL hot path with the reference's surface (ref:src/data/noisy_speech_dataset.py).

What changed relative to the reference, and why: the reference mixes and normalises every utterance on a DataLoader
worker (``__getitem__``: ~20 tensor ops + numpy z-norm per item).  Here the workers only do I/O -- load, mono,
crop/pad (ref:src/utils/audio_utils.py:9-62) and draw the noise file / SNR index -- and hand back RAW crops; the
main process copies the batch to the GPU (pinned, non-blocking) and ``GpuBatchMixer`` runs the fused
mix + peak-norm + z-norm kernel once per batch.  The batches the training loop sees are the reference's:
``{"clean_input_values": [B,1,L] f32, "noisy_input_values": [B,1,L] f32, "snr": [B] int64}``
(ref:src/data/noisy_speech_dataset.py:140-144), already on the device.

The reference's retry policy (up to 5 re-draws when ``add_noise_to_speech`` / normalisation rejects an item,
:56-148) is kept: the kernel reports one status word per row, the mixer re-draws the noise of rejected rows from the
other rows of the batch, and rows that still fail after ``max_attempts`` are replaced by the nearest good row of the batch (``GpuBatchMixer``).
"""
from __future__ import annotations

import os
import random
import wave
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, random_split

from .. import ops
from ..utils.logging_utils import logger

AUDIO_EXTENSIONS = {".wav", ".flac", ".mp3"}


def get_audio_files(directory: str) -> List[str]:
    """ref:src/utils/audio_utils.py:65-72 (sorted, so that the seeded split is reproducible across file systems)."""
    found = []
    for root, _, files in os.walk(directory):
        for f in files:
            if os.path.splitext(f)[1].lower() in AUDIO_EXTENSIONS:
                found.append(os.path.join(root, f))
    return sorted(found)


def _read_audio(path: str):
    """-> (waveform [C, N] float32, sample_rate).  torchaudio when it works, stdlib ``wave`` for PCM WAV otherwise
    (torchaudio.load needs TorchCodec, which is absent from this image: SURVEY.md fact 6)."""
    try:
        import torchaudio
        wav, sr = torchaudio.load(path)
        return wav.float(), int(sr)
    except Exception:
        pass
    with wave.open(path, "rb") as w:
        sr, ch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported WAV sample width {width} in {path}")
    return torch.from_numpy(data.reshape(-1, ch).T.copy()), int(sr)


def load_and_process_audio(file_path: str, sample_rate: int = 16000, max_audio_length: float = 5.0,
                           random_crop: bool = True) -> Optional[torch.Tensor]:
    """Mono, resampled, cropped / zero-padded waveform [1, max_samples], or ``None`` for unusable audio
    (same validity rules as ref:src/utils/audio_utils.py:9-62)."""
    try:
        max_samples = int(max_audio_length * sample_rate)
        wav, sr = _read_audio(file_path)
        if wav.shape[0] > 1:
            wav = wav.mean(dim=0, keepdim=True)
        if sr != sample_rate:
            import torchaudio
            wav = torchaudio.functional.resample(wav, sr, sample_rate)
        n = wav.shape[1]
        if n > max_samples:
            start = random.randint(0, n - max_samples) if random_crop else 0
            wav = wav[:, start:start + max_samples]
        elif n < max_samples:
            wav = torch.nn.functional.pad(wav, (0, max_samples - n))
        if torch.isnan(wav).any() or float(wav.abs().max()) < 1e-8:
            logger.warning("unusable audio (NaN or silent): %s", file_path)
            return None
        return wav.contiguous()
    except Exception as e:  # noqa: BLE001 -- the reference logs and skips any loader failure
        logger.error("error loading audio file %s: %s", file_path, e)
        return None


class NoiseRobustSpeechDataset(Dataset):
    """Same constructor as the reference (ref:src/data/noisy_speech_dataset.py:12-41).  Items are RAW crops:
    ``{"clean_wave": [1,L], "noise_wave": [1,L], "snr_idx": int, "snr": int}``; mixing happens on the GPU."""

    def __init__(self, clean_data_path: str, noise_data_path: str, sample_rate: int = 16000,
                 max_audio_length: float = 5.0, snr_range: Sequence[int] = (0, 5, 10, 15, 20),
                 feature_extractor=None):
        self.sample_rate = sample_rate
        self.max_samples = int(max_audio_length * sample_rate)
        self.snr_range = list(snr_range)
        self.feature_extractor = feature_extractor  # kept for interface parity; z-norm is fused into the mix kernel
        if feature_extractor is not None and not getattr(feature_extractor, "do_normalize", True):
            raise ValueError("the fused mix kernel implements do_normalize=True (wavlm preprocessor setting)")
        self.clean_files = get_audio_files(clean_data_path)
        self.noise_files = get_audio_files(noise_data_path)
        logger.info("Found %d clean files and %d noise files.", len(self.clean_files), len(self.noise_files))

    def __len__(self) -> int:
        return len(self.clean_files)

    def _load(self, path: str) -> Optional[torch.Tensor]:
        return load_and_process_audio(path, self.sample_rate, self.max_samples / self.sample_rate, random_crop=True)

    def __getitem__(self, idx: int) -> Dict[str, object]:
        max_attempts = 5  # ref:src/data/noisy_speech_dataset.py:56
        for attempt in range(max_attempts):
            clean = self._load(self.clean_files[idx])
            if clean is None:
                idx = (idx + 1) % len(self.clean_files)
                continue
            noise = self._load(self.noise_files[random.randint(0, len(self.noise_files) - 1)])
            if noise is None:
                continue
            snr_idx = random.randrange(len(self.snr_range))
            return {"clean_wave": clean, "noise_wave": noise, "snr_idx": snr_idx, "snr": int(self.snr_range[snr_idx])}
        raise RuntimeError(f"no usable audio after {max_attempts} attempts starting at index {idx}")


class TensorPairDataset(Dataset):
    """In-memory clean/noise waveforms with SNR cycled by index -- the recipe of ref:test/create_mock_dataset.py
    (clean randn, noise randn, snr = snr_range[i % n]) at real utterance lengths."""

    def __init__(self, clean: torch.Tensor, noise: torch.Tensor, snr_range: Sequence[int]):
        assert clean.dim() == 2 and noise.dim() == 2 and clean.shape[0] == noise.shape[0]
        self.clean, self.noise, self.snr_range = clean.float(), noise.float(), list(snr_range)

    def __len__(self) -> int:
        return self.clean.shape[0]

    def __getitem__(self, i: int):
        k = i % len(self.snr_range)
        return {"clean_wave": self.clean[i:i + 1], "noise_wave": self.noise[i:i + 1], "snr_idx": k,
                "snr": int(self.snr_range[k])}


class GpuBatchMixer:
    """Raw batch -> reference-format batch on the device: one fused mix + peak-norm + z-norm launch per batch, and NO
    host synchronisation.

    The reference's ``__getitem__`` retries an item up to five times when ``add_noise_to_speech`` or the peak checks
    reject it, drawing another noise file AND another SNR each time (ref:src/data/noisy_speech_dataset.py:55-149), and
    never emits an unusable item.  Here the decisions stay on the device:

    * after the first launch ONE retry launch redoes exactly the rows whose status is non-zero, looping INSIDE the kernel
      over up to ``max_attempts - 1`` further attempts with the following rows' noise crops and SNR draws
      (``ops.mix_batch`` = ``nrse_mix_batch_f32``); for a healthy batch its CTAs exit at once;
    * rows that are still bad after that (a silent or NaN CLEAN crop can never recover) are handled by ``bad_rows``:
      ``"substitute"`` (default) -- the finishing launch (which also gathers the ``snr`` labels and counts the rejected rows)
      copies the nearest following good row over them, the device-side form of the reference's "move on to the next item"
      (:60-66): the batch keeps its size, no all-zero waveform reaches BatchNorm statistics or the loss mean, and the
      host still never waits; ``"drop"`` -- read the status and remove the rows (one host synchronisation per batch);
      ``"keep"`` -- leave them zero-filled (BYOL mode) / as the clean waveform (emotion mode).
    * ``batch["snr"]`` follows the SNR index every row was finally mixed at; ``batch["mix_status"]`` keeps the failure
      code of every row that was ever rejected for good, and a count of such rows is copied to pinned memory
      asynchronously and logged when the NEXT batch is prepared (by then it has long arrived).

    With a single-row batch there is no other row to borrow noise or outputs from: such a row stays as the kernel left it.
    """

    def __init__(self, snr_range: Sequence[int], device, peak_norm: bool = True, max_attempts: int = 5,
                 bad_rows: str = "substitute", drop_bad_rows: Optional[bool] = None):
        if drop_bad_rows is not None:  # older spelling
            bad_rows = "drop" if drop_bad_rows else "keep"
        if bad_rows not in ("substitute", "drop", "keep"):
            raise ValueError("bad_rows must be 'substitute', 'drop' or 'keep'")
        self.snr_table = [float(v) for v in snr_range]
        self.device = torch.device(device)
        self.peak_norm = peak_norm
        self.max_attempts = max_attempts
        self.bad_rows = bad_rows
        self.rejected_rows = 0          # rows that stayed bad after all attempts, as far as already observed
        # rejected-row counts travel through a fixed ring of pinned int32 slots allocated ONCE (cudaHostAlloc inside a
        # training loop stalls the launch queue for milliseconds): slot i of batches [tail, head) is in flight
        self._ring = None               # pinned int32 [RING]
        self._ring_events = None
        self._ring_srcs = None          # the device counters, referenced until their copy has been seen
        self._head = self._tail = 0
        self._side = None               # stream of the count read-back: keeps the D2H copy off the step's critical path
        self._produced = None
        self._snr_labels = None

    RING = 64  # batches whose count may be in flight; a host that runs further ahead than this waits for the oldest one

    @property
    def drop_bad_rows(self) -> bool:
        return self.bad_rows == "drop"

    def _poll(self, block: bool = False) -> None:
        """Look at the counts that have arrived (in order; ``block``: wait for all outstanding ones)."""
        while self._tail < self._head:
            i = self._tail % self.RING
            ev = self._ring_events[i]
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            n = int(self._ring[i])
            if n:
                self.rejected_rows += n
                logger.error("%d row(s) failed all %d mix attempts (policy: %s)", n, self.max_attempts, self.bad_rows)
            self._ring_srcs[i] = None  # releases the device counter to the allocator
            self._tail += 1

    def _read_back(self, src: torch.Tensor) -> None:
        """Queue the asynchronous copy of a device int32 [1] count into the next ring slot, on the side stream."""
        if self._ring is None:
            self._ring = torch.zeros(self.RING, dtype=torch.int32).pin_memory()
            self._ring_events = [torch.cuda.Event() for _ in range(self.RING)]
            self._ring_srcs = [None] * self.RING
            self._side, self._produced = torch.cuda.Stream(device=self.device), torch.cuda.Event()
        if self._head - self._tail == self.RING:   # ring full: the oldest copy was queued 64 batches ago
            self._ring_events[self._tail % self.RING].synchronize()
            self._poll()
        i = self._head % self.RING
        # on the compute stream the 4-byte copy sat between the mix and the first frontend kernel (~12 us per step of
        # copy-engine latency); `src` stays referenced until the copy has been seen
        self._produced.record()
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._produced)
            self._ring[i:i + 1].copy_(src, non_blocking=True)
            self._ring_events[i].record()
        self._ring_srcs[i] = src
        self._head += 1

    def flush(self) -> int:
        """Wait for the outstanding status counts (end of an epoch); returns the total number of rejected rows."""
        self._poll(block=True)
        return self.rejected_rows

    @torch.no_grad()
    def __call__(self, raw: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        dev = self.device
        self._poll()
        clean = raw["clean_wave"].to(dev, non_blocking=True).flatten(1).contiguous().float()
        noise = raw["noise_wave"].to(dev, non_blocking=True).flatten(1).contiguous().float()
        snr_idx = torch.as_tensor(raw["snr_idx"]).to(dev, non_blocking=True).to(torch.int32).contiguous()
        retried = self.peak_norm and clean.shape[0] > 1 and self.max_attempts > 1
        if retried:
            # mix, device-side retries (one launch: rejected rows loop inside it over the following rows' noise crops and SNR
            # draws; a no-op on a healthy batch), substitute + labels + rejected-row count: ops.mix_batch
            if self._snr_labels is None or self._snr_labels.device != clean.device:
                # the label every row was FINALLY mixed at (a retry re-draws the SNR, as the reference's next attempt does)
                self._snr_labels = torch.tensor([int(round(v)) if float(v).is_integer() else v for v in self.snr_table],
                                                device=clean.device).to(torch.int64)
            c, n, status, _, snr, n_bad = ops.mix_batch(clean, noise, snr_idx, self.snr_table, self._snr_labels, True,
                                                        self.max_attempts, self.bad_rows == "substitute")
        else:
            c, n, status = ops.mix_normalize(clean, noise, snr_idx, self.snr_table, self.peak_norm)
            snr = torch.as_tensor(raw["snr"]).to(dev, non_blocking=True).to(torch.int64)
            n_bad = None
        if self.bad_rows == "drop":
            bad = status != 0
            if bool(bad.any()):  # host synchronisation
                keep = (~bad).nonzero().flatten()
                logger.error("dropping %d row(s) that failed %d mix attempts", int(bad.sum()), self.max_attempts)
                self.rejected_rows += int(bad.sum())
                c, n, snr, status = (c[keep] if c is not None else None), n[keep], snr[keep], status[keep]
        elif self.peak_norm:
            self._read_back(n_bad if n_bad is not None else (status != 0).sum().reshape(1).to(torch.int32))
        out = {"noisy_input_values": n.unsqueeze(1), "snr": snr, "mix_status": status}
        if self.peak_norm:
            out["clean_input_values"] = c.unsqueeze(1)
        return out


class MixedBatchLoader:
    """Iterates a raw DataLoader and yields device batches in the reference's format (len / dataset forwarded)."""

    def __init__(self, loader: DataLoader, mixer: GpuBatchMixer):
        self.loader, self.mixer = loader, mixer
        self.dataset = loader.dataset
        self.batch_size = loader.batch_size

    def __len__(self) -> int:
        return len(self.loader)

    def set_epoch(self, epoch: int) -> None:
        """Forwarded to the ``DistributedSampler`` (if any) so that every epoch is shuffled differently across ranks."""
        sampler = getattr(self.loader, "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        for raw in self.loader:
            yield self.mixer(raw)


def create_dataloaders(config, feature_extractor=None, device=None, dataset: Optional[Dataset] = None):
    """(train_loader, val_loader) as in ref:src/data/noisy_speech_dataset.py:151-194: seeded ``random_split``,
    ``pin_memory``, shuffle on train only.  ``device`` defaults to ``config['device']`` / cuda:0; ``dataset`` lets
    callers substitute an in-memory dataset (tests, benchmarks)."""
    data, training = config["data"], config["training"]
    if dataset is None:
        dataset = NoiseRobustSpeechDataset(data["clean_data_path"], data["noise_data_path"], data["sample_rate"],
                                           data["max_audio_length"], data["snr_range"], feature_extractor)
    val_size = int(len(dataset) * data.get("validation_ratio", 0.1))
    train_size = len(dataset) - val_size
    logger.info("Splitting dataset: %d training samples, %d validation samples", train_size, val_size)
    train_ds, val_ds = random_split(dataset, [train_size, val_size],
                                    generator=torch.Generator().manual_seed(training.get("seed", 42)))
    device = device or config.get("device", "cuda:0")
    mixer = GpuBatchMixer(data["snr_range"], device)
    kw = dict(batch_size=training["batch_size"], num_workers=training.get("num_workers", 4), pin_memory=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # data parallel: every rank iterates its own 1/world shard of both splits (``batch_size`` is per rank); call
        # ``train_loader.set_epoch(epoch)`` each epoch.  Without this every rank would draw the SAME batches (same seed)
        # and the all-reduce would average N copies of one gradient.
        from torch.utils.data.distributed import DistributedSampler
        seed = training.get("seed", 42)
        train = DataLoader(train_ds, sampler=DistributedSampler(train_ds, shuffle=True, seed=seed), **kw)
        val = DataLoader(val_ds, sampler=DistributedSampler(val_ds, shuffle=False, seed=seed), **kw)
    else:
        train = DataLoader(train_ds, shuffle=True, **kw)
        val = DataLoader(val_ds, shuffle=False, **kw)
    return MixedBatchLoader(train, mixer), MixedBatchLoader(val, mixer)
