#!/usr/bin/env python
"""Benchmark of the B200 BYOL noisy-view hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one pass of the hot path over one synthetic batch per GPU (64 utterances x 4 s @ 16 kHz):
fused SNR mix + peak-norm + z-norm of the clean/noisy views, then the WavLM-large conv feature encoder
forward on BOTH views (online = clean, target = noisy).  Metric: utterance-seconds per second, whole job.
`value` is timed with the inputs resident in HBM; `e2e` goes through the public module API with pinned HOST
buffers (H2D of the raw waveforms and D2H of the result inside the timed region).  `roofline` is measured live
with CUDA events around the dominant kernel (the tcgen05 implicit-GEMM conv layers); `cpu_baseline` times the
oracle (a CPU port of the reference path) on a bounded sample on rank 0.  `--impl reference` runs only that CPU
path and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 64
N_SAMPLES = 64000
SAMPLE_RATE = 16000
UTT_SEC_PER_STEP = BATCH * N_SAMPLES / SAMPLE_RATE  # 256 utterance-seconds per GPU per step
METRIC = "utterance-sec/sec, BYOL noisy-view step (SNR mix + WavLM-large conv frontend fwd, both views)"
UNIT = "utterance-seconds/s"
CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons / power every 50 ms while a timed region runs.  The sampler process is started
    (and has delivered its first sample) BEFORE the warm-up steps, so that its start-up -- NVML initialisation takes the
    driver's locks for a while -- does not fall into the timed region; samples are time-stamped and only those taken
    between ``mark_begin()`` and ``stop()`` are reported."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = f"/tmp/nrse_clocks_{os.getpid()}.csv"
        self.t_begin = None

    def start(self, wait_first_sample_s: float = 3.0):
        if os.environ.get("NRSE_BENCH_NO_SAMPLER"):  # diagnosis only: does the sampler process perturb the launches?
            return
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
            return
        t0 = time.time()
        while time.time() - t0 < wait_first_sample_s:  # nvidia-smi is up and sampling before anything is timed
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            if self.proc.poll() is not None:
                break
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.time()

    @staticmethod
    def parse_line(line: str):
        """One csv line -> (unix time or None, sm MHz, max sm MHz, power W, [active reason names]) or None."""
        f = [x.strip() for x in line.split(",")]
        if len(f) < 10:
            return None
        try:
            sm, smax, power = float(f[2]), float(f[3]), float(f[4])
        except ValueError:
            return None
        ts = None
        for fmt in ("%Y/%m/%d %H:%M:%S.%f", "%Y/%m/%d %H:%M:%S"):
            try:
                ts = datetime.datetime.strptime(f[0], fmt).timestamp()
                break
            except ValueError:
                continue
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [name for name, val in zip(names, f[6:10]) if val.lower().startswith("active")]
        return ts, sm, smax, power, reasons

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.out.close()
        rows = [r for r in (self.parse_line(line) for line in open(self.path)) if r is not None]
        try:
            os.remove(self.path)
        except OSError:
            pass
        inside = [r for r in rows if r[0] is None or self.t_begin is None or self.t_begin - 0.05 <= r[0] <= t_end + 0.05]
        used = inside if inside else rows[-1:]  # a region shorter than the sampling period: the closest sample
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = sorted({name for r in used for name in r[4]})
        return {"sm_mhz": statistics.median(r[1] for r in used), "sm_max_mhz": max(r[2] for r in used), "reasons": reasons,
                "power_w_max": max(r[3] for r in used), "samples": len(used)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU port of the reference path) on a bounded sample
# ------------------------------------------------------------------------------------------------------------------
CPU_SAMPLE_UTTS = 16  # of the 64 utterances of a step: the same bounded sample in `cpu_baseline` and `--impl reference`


def bench_config(world: int) -> dict:
    """The `config` object of the JSON line: identical in both arms (the CPU arm's bounded sample is described in its
    `cpu_baseline.sample`, not here)."""
    return {"workload": "configs[1]: SNR mix + WavLM-large conv frontend fwd, batch 64 x 4 s 16 kHz per GPU, "
                        "both BYOL views (clean->online, noisy->target)",
            "batch_per_gpu": BATCH, "n_samples": N_SAMPLES, "views": 2, "norm": "layer",
            "weights": "random init, WavLM-large conv shapes", "parallelism": f"dp{world} (no data-path collective)",
            "l2": "per-step working set 3.4 GB >> 126 MB L2: inputs are evicted between timed iterations"}


def cpu_reference_step(sample_utts: int, clean, noise, snr_idx, snr_table, layers):
    """One bounded-sample step of the SAME workload on the host cores: mix+normalise `sample_utts` utterances the
    way the reference's DataLoader worker does, then the fp32 conv feature encoder on both views."""
    import oracle
    t0 = time.perf_counter()
    c, n, st = oracle.mix_normalize_batch(clean[:sample_utts], noise[:sample_utts], snr_idx[:sample_utts], snr_table,
                                          peak_norm=True)
    with torch.no_grad():
        y_c = oracle.conv_frontend(c, layers, "layer")
        y_n = oracle.conv_frontend(n, layers, "layer")
    dt = time.perf_counter() - t0
    return dt, float(y_c.abs().mean() + y_n.abs().mean())


def run_cpu(sample_utts: int, steps: int, warmup: int):
    from nrse_b200.utils import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clean, noise, snr_idx, snr_table = synthetic.waveforms(sample_utts, N_SAMPLES, seed=1234)
    layers = synthetic.frontend_weights("layer", seed=0)
    for _ in range(warmup):
        cpu_reference_step(sample_utts, clean, noise, snr_idx, snr_table, layers)
    times = [cpu_reference_step(sample_utts, clean, noise, snr_idx, snr_table, layers)[0] for _ in range(steps)]
    dt = sum(times) / len(times)
    value = sample_utts * N_SAMPLES / SAMPLE_RATE / dt
    return value, dt, cores, torch.get_num_threads()


def cpu_sample_text(steps: int, warmup: int, threads: int, cores: int) -> str:
    return (f"{CPU_SAMPLE_UTTS} of {BATCH} utterances x 4 s per step, {steps} steps after {warmup} warm-up: oracle (CPU port "
            f"of the reference path: per-utterance mix+normalise, fp32 conv frontend on both views), torch {threads} "
            f"threads on {cores} host cores")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # each step is a bounded sample (16 of 64 utterances, ~1 s of CPU work): K steps after W warm-ups, as asked for,
    # capped so that the run ends within a few minutes
    steps = max(1, min(args.steps, 60))
    warmup = max(1, min(args.warmup, 5))
    value, dt, cores, threads = run_cpu(CPU_SAMPLE_UTTS, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": bench_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(steps, warmup, threads, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def load_frontend(module, layers, dev):
    """Synthetic WavLM-large-shaped conv weights into a B200FeatureEncoder (random init: no checkpoints offline)."""
    with torch.no_grad():
        for conv_layer, l in zip(module.conv_layers, layers):
            conv_layer.conv.weight.copy_(torch.from_numpy(l["conv"]))
            conv_layer.layer_norm.weight.copy_(torch.from_numpy(l["gamma"]))
            conv_layer.layer_norm.bias.copy_(torch.from_numpy(l["beta"]))
    return module.to(dev)


# launches of THIS repository's kernels per step (torch's own small kernels -- index / sum / mean -- are not counted)
MIXER_LAUNCHES = 1 + 1 + 1       # GpuBatchMixer (nrse_mix_batch_f32): mix, one retry launch (all further attempts loop inside
                                 # it; a no-op on a healthy batch), finish (substitute + snr labels + rejected-row count)
FRONTEND_FWD_LAUNCHES = 7        # layer 0 + six tcgen05 GEMM layers
FRONTEND_BWD_LAUNCHES = 7 + 6 + 12 + 1   # norm+GELU backward x7, wgrad x6, dgrad (even / odd) x6, layer-0 wgrad
PACK_LAUNCHES = 6                # bf16 re-pack of conv weights 1..6 after a parameter update


def main_gpu(args):
    import torch.distributed as dist

    from nrse_b200 import _lib, ops
    from nrse_b200.data import DevicePrefetcher, GpuBatchMixer
    from nrse_b200.models import B200FeatureEncoder, wavlm_large_config
    from nrse_b200.utils import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    for var in ("NRSE_EXPERIMENT", "NRSE_B200_LIB"):  # timing-experiment hooks of scripts/: never in a measured number
        if os.environ.get(var):
            raise SystemExit(f"bench.py refuses to run with {var} set (experiment hook: wrong results / another build)")
    if _lib.load().nrse_experiments_build() != 0:
        raise SystemExit("bench.py refuses the -DNRSE_EXPERIMENTS build of the library")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()

    # ---- synthetic inputs and WavLM-large-shaped frontend weights (random init: no checkpoints offline) ----------
    clean_np, noise_np, snr_idx_np, snr_table = synthetic.waveforms(BATCH, N_SAMPLES, seed=1234 + rank)
    layers = synthetic.frontend_weights("layer", seed=0)
    snr_list = [float(v) for v in snr_table]
    raw_h = {"clean_wave": torch.from_numpy(clean_np).pin_memory(), "noise_wave": torch.from_numpy(noise_np).pin_memory(),
             "snr_idx": torch.from_numpy(snr_idx_np).pin_memory(),
             "snr": torch.from_numpy(snr_table[snr_idx_np].astype(np.int64)).pin_memory()}
    raw_d = {k: v.to(dev) for k, v in raw_h.items()}
    clean_d, noise_d, snr_d = raw_d["clean_wave"], raw_d["noise_wave"], raw_d["snr_idx"]

    if args.train_only:  # tuning runs of the training-step leg alone
        def timed_only(fn, steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = fn()
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), out
        res = train_step_leg(args, dev, world, rank, layers, raw_d, snr_table, timed_only)
        if rank == 0:
            print(json.dumps({"train_step": res}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- the module surface a user of the reference calls (INTEGRATION.md level 1) -------------------------------------
    # GpuBatchMixer = the GPU half of NoiseRobustSpeechDataset.__getitem__ (mix + peak-norm + z-norm with the reference's
    # retry policy on the device); B200FeatureEncoder = WavLMModel.feature_extractor.  One encoder serves both views here:
    # online and target frontends have the same shapes and cost.
    mixer = GpuBatchMixer(snr_table.tolist(), dev)
    encoder = load_frontend(B200FeatureEncoder(wavlm_large_config()), layers, dev).eval()
    conv_w = [l.conv.weight.detach() for l in encoder.conv_layers]
    gammas = [l.layer_norm.weight.detach() for l in encoder.conv_layers]
    betas = [l.layer_norm.bias.detach() for l in encoder.conv_layers]
    packed = encoder._packed_weights()
    T, P = ops.frontend_geometry(N_SAMPLES)

    @torch.no_grad()
    def hot_path(raw):
        batch = mixer(raw)
        y_online = encoder(batch["clean_input_values"])   # [B, 512, T] fp32 (HF layout: a view of the kernels' [B, T, 512])
        y_target = encoder(batch["noisy_input_values"])
        return y_online, y_target, batch["mix_status"]

    launches_per_step = MIXER_LAUNCHES + 2 * FRONTEND_FWD_LAUNCHES

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    debug = bool(os.environ.get("NRSE_BENCH_DEBUG"))

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = []
        e0.record()
        for _ in range(steps):
            h0 = time.perf_counter()
            out = fn()
            host.append(time.perf_counter() - h0)
        if finish is not None:
            finish()  # work the loop left on other streams joins the timing stream before the closing event
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if debug and rank == 0:  # host enqueue time per step: a stalled launch queue shows up as one long step
            top = sorted(range(steps), key=lambda i: -host[i])[:3]
            print(f"[bench debug] {steps} steps, device {ms:.2f} ms, host enqueue sum {sum(host) * 1e3:.2f} ms, longest steps "
                  + ", ".join(f"#{i}: {host[i] * 1e3:.2f} ms" for i in top), file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # ---- value: inputs resident in HBM ----------------------------------------------------------------------------
    step_dev = lambda: hot_path(raw_d)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # up and sampling before the warm-up: its start-up must not perturb the timed region
    warmup = max(args.warmup, 3)
    out = None
    for _ in range(warmup):
        out = step_dev()  # the result is kept exactly as in the timed loop (the previous step's outputs are still alive while
        # a step allocates), and dropped before the timed region starts: the caching allocator then owns every block the
        # loop will ask for.  Without this the SECOND timed step was the first to need a third generation of output blocks
        # and called cudaMalloc next to running kernels -- 0.8 ms at best, 30-115 ms in one of six runs, inside a 50 ms
        # window (`NRSE_BENCH_DEBUG=1` prints the host time of the slowest steps; gpurun_out/r2x_*, r2w_*).
    out = None
    sampler.mark_begin()
    ms_total, out = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    assert int(out[2].abs().sum().item()) == 0, "synthetic batch produced rejected rows"
    ms_per_step = ms_total / args.steps
    value = world * UTT_SEC_PER_STEP / (ms_per_step * 1e-3)

    # ---- the same loop for >= 2 s: the sustained (power-capped) figure next to the short-run one ----------------------
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / ms_per_step) + 1)
        sus_sampler = ClockSampler(local_rank)
        if rank == 0:
            sus_sampler.start()
        sus_sampler.mark_begin()
        out = None
        ms_sus, _ = timed(step_dev, n_sus)
        sus_clocks = sus_sampler.stop() if rank == 0 else None
        sustained = {"value": world * UTT_SEC_PER_STEP / (ms_sus / n_sus * 1e-3), "unit": UNIT, "steps": n_sus,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus, "clocks": sus_clocks}

    # ---- e2e: the same module calls on pinned HOST buffers, H2D + D2H inside the timed region ----------------------------
    pooled_h = torch.empty(2, 2, BATCH, 512, dtype=torch.float32).pin_memory()   # [result slot, view, B, 512]
    status_h = torch.empty(2, BATCH, dtype=torch.int32).pin_memory()
    prefetch = DevicePrefetcher(dev, depth=2)
    prefetch.put(raw_h)  # pipeline prologue: the first batch is in flight before step 0

    done = [torch.cuda.Event(), torch.cuda.Event()]
    produced = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_stream = torch.cuda.Stream(device=dev)
    keep = [None, None]  # device sources of the result copies in flight (allocated on the compute stream, read on d2h_stream)
    n_e2e = [0]

    def step_e2e():
        # every step runs the hot path on the batch copied one step earlier, enqueues ONE H2D copy (the next step's raw
        # waveforms, on the copy stream: it overlaps this step's kernels) and copies its result to the host; the host then
        # waits for the PREVIOUS step's result (the one-step lag every training loop gives `loss.item()`): this step's
        # kernels are already queued, so the GPU never idles while the host prepares the next launch
        b = prefetch.get()
        y_o, y_t, st = hot_path(b)
        k = n_e2e[0] & 1
        po, pt = y_o.mean(dim=2), y_t.mean(dim=2)  # the [B,H]-pooled result the BYOL heads consume
        # the three small D2H copies run on their own stream behind an event: queued on the compute stream they sat between
        # this step's last kernel and the next step's first one (~14 us of copy-engine latency each).  Slot k's sources were
        # last read two steps ago, and that copy was waited for in the previous step, so `keep[k]` may be replaced.
        produced[k].record()
        keep[k] = (po, pt, st)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(produced[k])
            pooled_h[k, 0].copy_(po, non_blocking=True)
            pooled_h[k, 1].copy_(pt, non_blocking=True)
            status_h[k].copy_(st, non_blocking=True)
            done[k].record()
        prefetch.release()
        prefetch.put(raw_h)
        if n_e2e[0] > 0:
            done[k ^ 1].synchronize()  # the caller reads a result every step (the previous step's)
        n_e2e[0] += 1
        return y_o, y_t, st

    out = None
    for _ in range(3):
        out = step_e2e()  # same liveness pattern as the timed loop (see the warm-up of `value`)
    out = None
    # the last step's result copies are inside the timed region: the compute stream waits for them before the closing event
    ms_e2e_total, _ = timed(step_e2e, args.steps,
                            finish=lambda: torch.cuda.current_stream().wait_event(done[(n_e2e[0] - 1) & 1]))
    e2e_value = world * UTT_SEC_PER_STEP / (ms_e2e_total / args.steps * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in raw_h.values())
    d2h = (pooled_h.numel() * 4 + status_h.numel() * 4) // 2

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (conv operands/activations, fp32 accumulate + fp32 LayerNorm/GELU); f32 (mix)",
        "data": "synthetic", "config": bench_config(world),
        "value_sustained": None if sustained is None else sustained["value"],
        "sustained": sustained,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e_total / args.steps,
                "how": "pinned host batch of RAW waveforms -> DevicePrefetcher (copy stream, depth 2: the H2D copy of step "
                       "i+1 overlaps the kernels of step i) -> GpuBatchMixer (mix + device-side retries + substitute) -> "
                       "B200FeatureEncoder.forward on both views (the module calls of INTEGRATION.md level 1, eager, no CUDA "
                       "graph) -> D2H of the pooled [2,B,512] features + status every step (copy stream behind an event; "
                       "the last step's copies complete inside the timed region); the host waits for the PREVIOUS step's "
                       "result after queueing the current one (one-step lag, as for loss.item() in a training loop)"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }

    # ---- the north-star TRAINING step on the hot path, with its collective ---------------------------------------------------
    if not args.no_train_step:
        line["train_step"] = train_step_leg(args, dev, world, rank, layers, raw_d, snr_table, timed)
        if world > 1 and args.train_allreduce == "full":
            # the same step with only the hot path's OWN trainable parameters (conv frontend + heads, 14 M) in the gradient
            # set: what the path costs under data parallelism when the transformer's 1.24 GB of gradients travel where
            # they do in the real step -- behind the transformer's backward, which this leg excludes
            import copy
            args2 = copy.copy(args)
            args2.train_allreduce = "hotpath"
            torch.cuda.empty_cache()
            line["train_step_hot_path_gradients_only"] = train_step_leg(args2, dev, world, rank, layers, raw_d, snr_table, timed)

    # ---- per-kernel rooflines (rank 0 only, outside the headline timing) -------------------------------------------
    if rank == 0:
        # the per-kernel numbers are "kernel timed alone" numbers: let the board leave the power-capped state of the long
        # timed regions above first, and record the clocks this section actually ran at
        torch.cuda.synchronize()
        time.sleep(2.0)
        ksampler = ClockSampler(local_rank)
        ksampler.start()
        line.update(kernel_rooflines(ops, dev, peaks, clean_d, noise_d, snr_d, snr_list, conv_w, gammas, betas, packed,
                                     T, P, max(3, min(args.steps, 10))))
        line["kernel_clocks"] = ksampler.stop()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cv, cdt, cores, threads = run_cpu(CPU_SAMPLE_UTTS, 4, 1)
        line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": cpu_sample_text(4, 1, threads, cores)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def wavlm_large_ballast_shapes():
    """Shapes of every WavLM-large parameter OUTSIDE the conv feature encoder (feature projection, positional conv, 24
    transformer layers: 311 M parameters), from the architecture on the meta device -- no memory, no weights."""
    from transformers import WavLMModel
    from nrse_b200.models import wavlm_large_config
    with torch.device("meta"):
        m = WavLMModel(wavlm_large_config())
    return [tuple(p.shape) for n, p in m.named_parameters() if not n.startswith("feature_extractor.")]


def train_step_leg(args, dev, world, rank, layers, raw_d, snr_table, timed):
    """BYOL TRAINING step restricted to the hot path of SURVEY.md section 8, data-parallel with the north star's collective
    (ref:train_byol.py:47-77): GpuBatchMixer -> tape-writing conv frontend on the clean view (online) + inference frontend
    on the noisy view (target, no grad) -> mean-pool -> stock projector / predictor heads (BatchNorm in train mode) ->
    byol_loss -> backward through the heads and the NATIVE frontend backward -> gradient all-reduce over NCCL ->
    FusedAdamWEma (clip + AdamW + EMA of the target twins, two launches).

    The 24-layer transformer's COMPUTE is excluded (it is stock HF code, out of scope); its 311 M parameters are present as
    what they cost on this path: their gradients (constant values), AdamW moments and EMA twins go through the all-reduce
    and the fused optimizer, so the step moves the full 1.3 GB gradient of WavLM-large BYOL across NVLink and 13 GB
    through the optimizer kernels.  The transformer bucket group is all-reduced right after the loss (in the real step those
    gradients are complete before the conv frontend's backward starts: the frontend is the first layer) and overlaps the
    heads' and the frontend's backward; the frontend + heads group follows the backward and is the un-overlappable tail."""
    import torch.distributed as dist

    from nrse_b200 import ops
    from nrse_b200.data import GpuBatchMixer
    from nrse_b200.models import B200FeatureEncoder, PredictionHead, ProjectionHead, byol_loss, wavlm_large_config
    from nrse_b200.train import FusedAdamWEma, GradArena

    torch.manual_seed(0)  # identical initial weights on every rank
    cfg = wavlm_large_config()
    online_fe = load_frontend(B200FeatureEncoder(cfg), layers, dev).train()
    target_fe = load_frontend(B200FeatureEncoder(cfg), layers, dev).train()
    online_proj, online_pred = ProjectionHead(512, 1024, 1024).to(dev).train(), PredictionHead(1024, 2048, 1024).to(dev).train()
    target_proj = ProjectionHead(512, 1024, 1024).to(dev).train()
    target_proj.load_state_dict(online_proj.state_dict())
    for p in (*target_fe.parameters(), *target_proj.parameters()):
        p.requires_grad = False
    hot = [*online_fe.parameters(), *online_proj.parameters(), *online_pred.parameters()]
    ballast, ballast_twin = [], []
    if args.train_allreduce == "full":
        for shape in wavlm_large_ballast_shapes():
            ballast.append(torch.nn.Parameter(torch.randn(shape, device=dev) * 0.02))
            ballast_twin.append(ballast[-1].detach().clone())
    arena = GradArena([ballast, hot] if ballast else [hot], bucket_bytes=args.train_bucket_mb << 20,
                      multimem={"auto": "auto", "multimem": True, "nccl": False}[args.train_allreduce_impl],
                      multimem_ctas=args.train_multimem_ctas)
    g_ballast, g_hot = (0, 1) if ballast else (None, 0)
    if ballast:
        arena.flat[:arena.group_ranges[0][1]].normal_(0.0, 1e-4)  # the transformer's gradients: constant synthetic values
    ema_pairs = list(zip(online_fe.parameters(), target_fe.parameters())) + \
        list(zip(online_proj.parameters(), target_proj.parameters())) + list(zip(ballast, ballast_twin))
    opt = FusedAdamWEma([*ballast, *hot], lr=1e-5, weight_decay=1e-5, max_grad_norm=1.0,   # ref:train_byol.py:143-150
                        ema_pairs=[(o, t.data) for o, t in ema_pairs], ema_decay=0.996)
    opt.keep_grads = True  # gradients live in the arena: addresses (and the optimizer's chunk table) never change
    mixer = GpuBatchMixer(snr_table.tolist(), dev)

    def step(sync: bool):
        batch = mixer(raw_d)
        arena.zero_(g_hot)
        emb = online_fe(batch["clean_input_values"]).mean(dim=2)            # [B, 512]
        pred = online_pred(online_proj(emb))
        with torch.no_grad():
            tproj = target_proj(target_fe(batch["noisy_input_values"]).mean(dim=2))
        loss = byol_loss(pred, tproj)
        overlap = sync and world > 1
        if overlap and g_ballast is not None:
            # the persistent conv kernels leave a few SMs to NCCL while the collective is in flight (a machine filled
            # with one resident CTA per SM would make NCCL's kernels wait for the end of every kernel)
            if args.train_sm_reserve:
                ops.set_sm_budget(148 - args.train_sm_reserve)
            arena.all_reduce_async(g_ballast)   # travels while the backward below computes
        loss.backward()
        if overlap:
            arena.all_reduce_async(g_hot)
            arena.wait()
            if args.train_sm_reserve:
                ops.set_sm_budget(148)
        opt.step()
        return loss

    steps = max(5, min(args.steps, 50))
    for _ in range(5):  # allocator (3.3 GB tape per step), optimizer tables, NCCL channels
        step(True)
    # no-sync / sync / no-sync: the two no-sync loops bracket the measured one, so a drift of the board's power state
    # does not end up in `allreduce_exposed_ms`
    ms_a, _ = timed(lambda: step(False), steps)
    ms_sync, loss = timed(lambda: step(True), steps)
    ms_b, _ = timed(lambda: step(False), steps)
    ms_nosync = 0.5 * (ms_a + ms_b)
    assert bool(torch.isfinite(loss)), "training step produced a non-finite loss"
    n_hot, n_ballast = sum(p.numel() for p in hot), sum(p.numel() for p in ballast)
    n_coll = sum(len(arena.buckets(g)) for g in range(len(arena.groups))) if world > 1 else 0
    launches = MIXER_LAUNCHES + 2 * FRONTEND_FWD_LAUNCHES + 2 + FRONTEND_BWD_LAUNCHES + 3 * PACK_LAUNCHES + 2
    return {
        "what": "BYOL training step on the hot path: mix -> train fwd (online) + fwd (target) -> pool -> stock heads -> "
                "byol_loss -> heads + native frontend backward -> NCCL gradient all-reduce -> fused clip+AdamW+EMA; "
                "transformer compute excluded" + (", its 311 M parameters kept as gradient / optimizer / EMA / all-reduce "
                                                  "traffic" if ballast else " (and its parameters too: --train-allreduce hotpath)"),
        "value": world * UTT_SEC_PER_STEP / (ms_sync / steps * 1e-3), "unit": UNIT, "steps": steps,
        "ms_per_step": ms_sync / steps, "ms_per_step_no_allreduce": ms_nosync / steps,
        "allreduce_exposed_ms": (ms_sync - ms_nosync) / steps if world > 1 else 0.0,
        "ms_per_step_no_allreduce_runs": [ms_a / steps, ms_b / steps],
        "allreduce_bytes_per_rank": arena.numel * 4 if world > 1 else 0, "allreduce_collectives_per_step": n_coll,
        "allreduce_dtype": "f32", "allreduce_impl": arena.impl, "bucket_bytes": arena.bucket_elems * 4, "world": world,
        "sm_reserved_for_nccl": args.train_sm_reserve if world > 1 else 0,
        "trainable_params": n_hot + n_ballast, "hot_path_params": n_hot, "ema_params": sum(t.numel() for _, t in ema_pairs),
        "batch_per_gpu": BATCH, "n_samples": N_SAMPLES, "loss": float(loss.item()),
        "gpu_launches_per_step": launches,
    }


def kernel_rooflines(ops, dev, peaks, clean_d, noise_d, snr_d, snr_list, conv_w, gammas, betas, packed, T, P, reps):
    """CUDA-event timing of each hot-path kernel on its own launch stream; algorithmic work from SURVEY.md 8(d)."""
    def ev_time(fn, n=reps, warm=2):
        """Average device time of fn(): n calls captured in ONE CUDA graph (no host launch overhead between them),
        replayed once untimed and once between CUDA events on the launching stream."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(n):
                fn(i) if fn.__code__.co_argcount else fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n  # ms per call

    def plain_time(fn, n=5, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    B, L = clean_d.shape
    # the conv GEMM layers: run layer by layer on real activations so every launch can be bracketed by events
    c, n, _ = ops.mix_normalize(clean_d, noise_d, snr_d, snr_list, True)
    act = ops.conv_layer0(c, conv_w[0], gammas[0], betas[0], "layer").view(B * P[0], 512)
    t_l0 = ev_time(lambda: ops.conv_layer0(c, conv_w[0], gammas[0], betas[0], "layer"))
    gemm_ms, gemm_flops, per_layer = 0.0, 0.0, []
    for i in range(1, 7):
        k = CONV_KERNEL[i]
        inp = act
        # the per-layer entry point cannot know the layer: select what nrse_conv_frontend_fwd runs for it (default
        # variant 4: the 2-SM UMMA kernel for layers 1-3, the 1-SM CTA-pair kernel for layers 4-6)
        ops.set_frontend_variant(3 if i <= 3 else 2)
        t = ev_time(lambda: ops.conv_layer(inp, packed[i - 1], k, gammas[i], betas[i]))
        act = ops.conv_layer(inp, packed[i - 1], k, gammas[i], betas[i])
        ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)
        flops = 2.0 * B * T[i] * 512 * (512 * k)  # algorithmic: valid frames only
        gemm_ms += t
        gemm_flops += flops
        per_layer.append({"layer": i, "ms": t, "tflops": flops / (t * 1e-3) / 1e12})
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    # Two denominators, two timings.  (1) Each layer timed ALONE (10 graph-replayed launches, ~5 ms, after a 2 s cool-down:
    # the board runs at its burst clock) against the BURST cuBLAS peak -- the headline `frac`.  (2) The six launches of a
    # view back to back for >= 2 s (the board settles at its power cap, as in a real step) against the SUSTAINED peak.
    peak = peaks["bf16_tflops"]
    seq_graph = torch.cuda.CUDAGraph()
    seq_acts = [ops.conv_layer0(c, conv_w[0], gammas[0], betas[0], "layer").view(B * P[0], 512)]

    def run_six():
        a = seq_acts[0]
        for i in range(1, 7):
            ops.set_frontend_variant(3 if i <= 3 else 2)
            a = ops.conv_layer(a, packed[i - 1], CONV_KERNEL[i], gammas[i], betas[i])
        ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)
    run_six()
    torch.cuda.synchronize()
    with torch.cuda.graph(seq_graph):
        run_six()
    seq_graph.replay()
    torch.cuda.synchronize()
    n_rep = max(20, int(2200.0 / max(gemm_ms, 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_rep):
        seq_graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sus_ms = e0.elapsed_time(e1) / n_rep
    sus_achieved = gemm_flops / (sus_ms * 1e-3) / 1e12
    del seq_graph, seq_acts
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2z_ncu_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of the six launches, from the ncu capture
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["traffic_bytes_six_launches"], tj["source"]
    roofline = {"bound": "tensor", "kernel": "conv_gemm2_kernel (layers 1-3, 2-SM UMMA) + conv_gemm_kernel (layers 4-6): tcgen05 implicit GEMM + LayerNorm + GELU",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "bytes per 6 launches (one view)", "traffic_source": traffic_src,
                "peak_source": f"{peaks['source']} bf16_tflops (BURST cuBLAS peak: every layer is timed alone, graph-replayed "
                               "back to back for ~5 ms after a 2 s cool-down)",
                "launches": 6, "ms_per_view": gemm_ms, "algorithmic_flops_per_view": gemm_flops, "per_layer": per_layer,
                "sustained": {"achieved": sus_achieved, "peak": peaks["bf16_tflops_sustained"],
                              "frac": sus_achieved / peaks["bf16_tflops_sustained"], "ms_per_view": sus_ms,
                              "seconds": sus_ms * n_rep * 1e-3,
                              "how": "the six GEMM-layer launches of one view replayed back to back for >= 2 s (power-capped "
                                     "clocks) against the measured sustained cuBLAS peak"}}

    hbm = peaks["hbm_gbs"]
    kernels = []
    # the >= 2 s sustained replay above leaves the board power-capped (~1.45 GHz); the kernels below are timed ALONE against the
    # burst copy bandwidth of MEASURED_PEAKS.json, and the latency-bound ones among them (mix at B = 64) follow the SM clock:
    # let the clocks recover first (round 2 measured 24 us instead of 17 us for the same mix kernel without this pause)
    torch.cuda.synchronize()
    time.sleep(2.0)
    # 6 distinct input sets (6 x 65 MB in+out > 126 MB L2) cycled inside the graph: every launch reads from HBM
    sets = [(clean_d.clone(), noise_d.clone()) for _ in range(6)]
    t_mix = ev_time(lambda i=0: ops.mix_normalize(sets[i % 6][0], sets[i % 6][1], snr_d, snr_list, True), n=24)
    del sets
    kernels.append({"kernel": "mix_normalize (B=64 x 64000, inputs rotated through 6 buffer sets > L2)", "bound": "hbm", "ms": t_mix,
                    "achieved": 16.0 * B * L / (t_mix * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s"})
    big = 512
    cb = clean_d.repeat(big // B, 1).contiguous(); nb = noise_d.repeat(big // B, 1).contiguous(); sb = snr_d.repeat(big // B)
    t_mix_big = ev_time(lambda: ops.mix_normalize(cb, nb, sb, snr_list, True))
    kernels.append({"kernel": "mix_normalize (B=512, 524 MB > L2)", "bound": "hbm", "ms": t_mix_big,
                    "achieved": 16.0 * big * L / (t_mix_big * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s"})
    del cb, nb
    kernels.append({"kernel": "layer0_kernel (conv k=10 + LayerNorm + GELU, bf16 out)", "bound": "hbm", "ms": t_l0,
                    "achieved": (4.0 * B * L + 2.0 * 512 * B * T[0]) / (t_l0 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s"})
    # EMA over the WavLM-large encoder + projector parameter sizes (317,556,416 fp32 values, 12 B each)
    sizes = [8] * 24 + [16] * 24 + [128] + [512] * 40 + [1024] * 227 + [4096] * 24 + [5120] * 2 + [524288] * 3 + \
            [786432] * 4 + [1048576] * 98 + [4194304] * 48 + [8388608]
    online = [torch.randn(s, device=dev) for s in sizes]
    target = [torch.randn(s, device=dev) for s in sizes]
    plan = ops.EmaPlan(online, target)
    t_ema = ev_time(lambda: plan.step(0.996))
    kernels.append({"kernel": "ema_chunks_kernel (317.6 M params, 496 tensors, 1 launch)", "bound": "hbm", "ms": t_ema,
                    "achieved": 12.0 * plan.numel / (t_ema * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s"})
    del online, target, plan
    # fused optimizer tail (clip_grad_norm_ + AdamW + EMA, ref:train_byol.py:67-71) at WavLM-large size: 325,958,336
    # trainable parameters, of which 317,556,416 have an EMA twin; 4 B (norm read) + 28 B (p,g,m,v read; p,m,v write)
    # per parameter + 8 B (t read, t write) per twin
    from nrse_b200.train import FusedAdamWEma
    pred_sizes = [1024 * 2048, 2048, 2048, 2048, 2048 * 2048, 2048, 2048, 2048, 2048 * 1024, 1024]
    prm = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.02) for s in sizes + pred_sizes]
    twin = [torch.randn(s, device=dev) * 0.02 for s in sizes]
    for q in prm:
        q.grad = torch.randn_like(q) * 1e-3
    fopt = FusedAdamWEma(prm, lr=1e-5, weight_decay=1e-5, max_grad_norm=1.0, ema_pairs=zip(prm[:len(twin)], twin),
                         ema_decay=0.996)
    fopt.step()
    t_opt = ev_time(fopt.step, n=10)
    n_upd, n_tw = sum(q.numel() for q in prm), sum(q.numel() for q in twin)
    kernels.append({"kernel": "grad_sqnorm_chunks + adamw_ema_chunks (clip + AdamW + EMA, 325.9 M params, 2 launches)",
                    "bound": "hbm", "ms": t_opt, "achieved": (32.0 * n_upd + 8.0 * n_tw) / (t_opt * 1e-3) / 1e9,
                    "peak": hbm, "unit": "GB/s"})
    ref_opt = torch.optim.AdamW(prm, lr=1e-5, weight_decay=1e-5)  # ref:train_byol.py:146; foreach on CUDA

    def stock_tail():
        torch.nn.utils.clip_grad_norm_(prm, 1.0)
        ref_opt.step()
        for i in range(len(twin)):
            twin[i] = 0.996 * twin[i] + (1 - 0.996) * prm[i].data
    t_stock_tail = plain_time(stock_tail, n=3, warm=3)  # first calls build the AdamW state and settle the allocator
    del fopt, ref_opt, prm, twin
    torch.cuda.empty_cache()
    p = torch.randn(B, 1024, device=dev, requires_grad=True)
    z = torch.randn(B, 1024, device=dev)
    t_loss = ev_time(lambda: ops.byol_loss(p, z))
    kernels.append({"kernel": "byol_loss_fwd (64x1024 fp32, 1 launch, 0 syncs)", "bound": "hbm", "ms": t_loss,
                    "achieved": 8.0 * B * 1024 / (t_loss * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s"})
    for kk in kernels:
        kk["frac"] = kk["achieved"] / kk["peak"]

    # the same three operations written with stock torch ops on THIS GPU (what the reference's arithmetic costs when it
    # is simply moved to the device): batched mix + peak-norm + z-norm, the per-tensor EMA loop, the ~15-op loss
    def stock_mix(cw, nw, snr_lin):
        ps, pn = (cw ** 2).mean(1, keepdim=True), (nw ** 2).mean(1, keepdim=True)
        y = cw + nw * torch.sqrt(ps / (pn * snr_lin))
        outs = []
        for v in (cw, y):
            v = v / (v.abs().amax(1, keepdim=True) + 1e-8)
            outs.append((v - v.mean(1, keepdim=True)) / torch.sqrt(v.var(1, unbiased=False, keepdim=True) + 1e-7))
        return outs

    snr_lin = torch.tensor([10 ** (snr_list[i] / 10) for i in snr_d.tolist()], device=dev)[:, None]
    stock_ms = {"mix_ms": plain_time(lambda: stock_mix(clean_d, noise_d, snr_lin))}
    on = [torch.randn(s, device=dev) for s in sizes]
    tg = [torch.randn(s, device=dev) for s in sizes]

    def stock_ema():
        for i in range(len(on)):  # ref:src/models/byol.py:64-73, one tensor at a time
            tg[i] = 0.996 * tg[i] + (1 - 0.996) * on[i]
    stock_ms["ema_ms"] = plain_time(stock_ema)
    del on, tg

    def stock_loss():
        a = torch.nn.functional.normalize(p.detach() + 1e-10, dim=1, eps=1e-10)
        b_ = torch.nn.functional.normalize(z + 1e-10, dim=1, eps=1e-10)
        return 2 - 2 * torch.clamp((a * b_).sum(1), -1.0, 1.0).mean()
    stock_ms["loss_fwd_ms"] = plain_time(stock_loss)
    stock_ms["clip_adamw_ema_ms"] = t_stock_tail

    # training forward (tape-writing) + native backward of the frontend, and the same stack in stock torch on THIS GPU
    # (cuDNN conv1d + ATen layer_norm / gelu): the "kernel to beat" of BASELINE.md section 4 (G0)
    import torch.nn.functional as F
    x = c
    dpacks = [ops.pack_conv_weight_dgrad(w) for w in conv_w[1:]]
    gy = torch.randn(B, T[6], 512, device=dev)

    def plain(fn, n=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    y_t, tape = ops.conv_frontend_train(x, conv_w, gammas, betas, "layer", packed=packed)
    t_train = plain(lambda: ops.conv_frontend_train(x, conv_w, gammas, betas, "layer", packed=packed))
    t_bwd = plain(lambda: ops.conv_frontend_backward(x, conv_w, gammas, betas, tape, gy, "layer", dgrad_packs=dpacks))
    t_inf = plain(lambda: ops.conv_frontend(x, conv_w, gammas, betas, "layer", out_dtype=torch.bfloat16, packed=packed))
    ws = [w.clone().requires_grad_(True) for w in conv_w]
    gs = [w.clone().requires_grad_(True) for w in gammas]
    bs = [w.clone().requires_grad_(True) for w in betas]

    def stock(xx):
        h = xx[:, None]
        for i, wt in enumerate(ws):
            h = F.conv1d(h, wt, stride=ops.CONV_STRIDE[i])
            h = F.gelu(F.layer_norm(h.transpose(1, 2), (512,), gs[i], bs[i], 1e-5).transpose(1, 2))
        return h

    stock_res = {}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        def fo():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                return stock(x)

        def fb():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                out = stock(x)
            out.backward(gy.transpose(1, 2).to(out.dtype))
        stock_res[name] = {"fwd_ms": plain(fo), "fwd_bwd_ms": plain(fb)}
    frontend_train = {"shape": [B, L], "fwd_ms": t_inf, "train_fwd_ms": t_train, "bwd_ms": t_bwd,
                      "fwd_bwd_ms": t_train + t_bwd, "stock_torch_same_gpu": stock_res,
                      "note": "host-launched (not graph-timed): includes Python/launch overhead of the op wrappers"}
    return {"roofline": roofline, "kernels": kernels, "frontend_train": frontend_train,
            "stock_torch_same_gpu_ms": stock_ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain-seconds", type=float, default=2.5,
                    help="length of the second, sustained timed region of the device-resident loop (0 = skip)")
    ap.add_argument("--no-train-step", action="store_true", help="skip the hot-path training-step leg")
    ap.add_argument("--train-only", action="store_true", help="only the training-step leg (tuning runs; not a bench line)")
    ap.add_argument("--train-sm-reserve", type=int, default=32,
                    help="SMs the conv kernels leave to NCCL while the gradient all-reduce overlaps the backward (N > 1)")
    ap.add_argument("--train-bucket-mb", type=int, default=256, help="all-reduce bucket size of the training-step leg")
    ap.add_argument("--train-allreduce-impl", default="nccl", choices=["auto", "multimem", "nccl"],
                    help="gradient all-reduce of the training-step leg: this repository's NVSwitch-multicast kernel (auto: "
                         "when available) or NCCL all-reduce calls")
    ap.add_argument("--train-multimem-ctas", type=int, default=0, help="CTAs of the multicast all-reduce kernel (0 = 148)")
    ap.add_argument("--train-allreduce", default="full", choices=["full", "hotpath"],
                    help="gradient set of the training-step leg: WavLM-large BYOL's full 326 M parameters (default) or "
                         "only the hot path's own trainable parameters")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
