#!/usr/bin/env python
"""BASELINE.json configs[2]: the whole BYOL training step (clean/noisy views, loss, backward, clip, AdamW, EMA) on a
WavLM-large-shaped model (random init: no checkpoints offline), data-parallel with one process per GPU.

    python scripts/train_step_bench.py [--batch 64] [--seconds 4] [--steps 5] [--autocast]
    torchrun --nproc-per-node N scripts/train_step_bench.py ...

The hot-path kernels of this repository run inside it (GPU mix, conv frontend forward on both branches, fused loss,
fused clip + AdamW + EMA optimizer tail, native frontend backward); the 24-layer transformer and the heads are stock
PyTorch.  Prints one
JSON line per run with the step time, utterance-seconds/s, and the device time of the hot-path pieces.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nrse_b200.data import GpuBatchMixer  # noqa: E402
from nrse_b200.models import BYOLSpeechModel, wavlm_large_config  # noqa: E402
from nrse_b200.train import FusedAdamWEma, byol_step, init_distributed, wrap_data_parallel  # noqa: E402
from nrse_b200.utils import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=24, help="transformer layers (24 = wavlm-large)")
    ap.add_argument("--autocast", action="store_true", help="bf16 autocast for the stock transformer / heads")
    ap.add_argument("--layerdrop", type=float, default=0.1, help="WavLM LayerDrop (0.1 = wavlm-large; 0 for a "
                    "deterministic amount of work per step when A/B-ing)")
    ap.add_argument("--profile", action="store_true", help="also report the summed CUDA kernel time of one step "
                    "(torch.profiler): step time >> kernel time means the step is host-launch-bound")
    ap.add_argument("--keep-grads", action="store_true", help="fused optimizer: zero gradients in place (stable addresses)")
    ap.add_argument("--optimizer", choices=["fused", "torch"], default="fused",
                    help="fused: FusedAdamWEma (clip + AdamW + EMA in two launches); torch: the reference's sequence")
    args = ap.parse_args()

    rank, world, local_rank = init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    L = int(args.seconds * 16000)
    cfg = {"model": {"name": wavlm_large_config(num_hidden_layers=args.layers, layerdrop=args.layerdrop), "projection_dim": 1024,
                     "prediction_dim": 2048, "ema_decay": 0.997},
           "data": {"snr_range": [2, 5, 10, 15, 20]}}
    model = BYOLSpeechModel(cfg).to(dev)
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    ddp = wrap_data_parallel(model, dev)
    if args.optimizer == "fused":
        opt = FusedAdamWEma.for_byol(model, lr=1e-5, weight_decay=1e-5, max_grad_norm=1.0)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=1e-5)  # ref:train_byol.py:146
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=1000)
    if args.optimizer == "fused":
        opt.keep_grads = args.keep_grads
    clean, noise, snr_idx, table = synthetic.waveforms(args.batch, L, seed=1234 + rank)
    raw = {"clean_wave": torch.from_numpy(clean)[:, None].pin_memory(), "noise_wave": torch.from_numpy(noise)[:, None].pin_memory(),
           "snr_idx": torch.from_numpy(snr_idx), "snr": torch.tensor([table[i] for i in snr_idx])}
    mixer = GpuBatchMixer([2, 5, 10, 15, 20], dev)

    def step():
        batch = mixer(raw)  # H2D + fused mix/normalise
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.autocast):
            return byol_step(ddp, batch["clean_input_values"], batch["noisy_input_values"], opt, sched)

    ddp.train()
    for _ in range(args.warmup):
        loss = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    kernel_ms = top = None
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        evs = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
        kernel_ms = sum(e.device_time_total for e in evs) / 2 / 1e3
        top = [(e.key[:60], round(e.device_time_total / 2 / 1e3, 2), e.count // 2)
               for e in sorted(evs, key=lambda e: -e.device_time_total)[:12]]

    # device time of the hot-path pieces inside that step
    def ev(fn, n=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    batch = mixer(raw)
    fe = model.online_encoder.model.feature_extractor
    with torch.no_grad():
        t_fe = ev(lambda: fe(batch["clean_input_values"].squeeze(1)))
    t_ema = ev(model._update_target_network)
    t_mix = ev(lambda: mixer(raw))
    if rank == 0:
        print(json.dumps({
            "workload": "configs[2]: BYOL training step, WavLM-large shapes (random init), data-parallel",
            "n_gpus": world, "batch_per_gpu": args.batch, "seconds": args.seconds, "layers": args.layers,
            "autocast_bf16": args.autocast, "layerdrop": args.layerdrop, "optimizer": args.optimizer, "trainable_params": n_params,
            "loss": float(loss), "optimizer_table_builds": getattr(opt, "table_builds", None), "keep_grads": args.keep_grads,
            "ms_per_step": ms, "cuda_kernel_ms_per_step": kernel_ms, "top_kernels": top,
            "utterance_seconds_per_s": world * args.batch * args.seconds / (ms * 1e-3),
            "hot_path_ms": {"h2d+mix": t_mix, "conv_frontend_fwd_one_view": t_fe, "ema_update": t_ema},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
