#!/usr/bin/env python
"""BASELINE.json configs[3]: the emotion-dimension fine-tune step (default_wavlm-large_emotion_dim_ft.yaml:
batch 36 x 5 s, AdamW lr 5e-6 / wd 1e-4, noise added during training, encoder unfrozen) on a WavLM-large-shaped
EmotionClassifier (random init: no checkpoints offline), data-parallel with one process per GPU.

    python scripts/emotion_step_bench.py [--batch 36] [--seconds 5] [--steps 5] [--autocast]
    torchrun --nproc-per-node N scripts/emotion_step_bench.py ...

Inside the step: H2D of the raw crops, GPU mix in the emotion mode (a-3': mix + z-norm, no peak norm), the B200 conv
feature encoder forward + native backward, batched attentive statistics pooling (2 + 2 launches instead of the
reference's per-utterance loop), CCC loss, fused clip + AdamW.  The 24-layer transformer and the MLP heads are stock
PyTorch.  ``--stock-pool`` swaps in the reference-style per-utterance pooling loop (stock torch ops on the same GPU) and
``--optimizer torch`` the reference's clip + AdamW for an A/B of the two consumer-side pieces."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from nrse_b200.data import GpuBatchMixer  # noqa: E402
from nrse_b200.models import EmotionClassifier, WavLMEncoder, wavlm_large_config  # noqa: E402
from nrse_b200.train import FusedAdamWEma, emotion_dim_step, init_distributed, wrap_data_parallel  # noqa: E402
from nrse_b200.utils import synthetic  # noqa: E402


def stock_pool_forward(self, xs, mask):
    """The reference's loop (ref:src/models/pool.py:44-58) with stock torch ops, for the A/B."""
    feat_lens = self.compute_length_from_mask(mask).tolist()
    pooled = []
    for x, n in zip(xs, feat_lens):
        x = x[:n].unsqueeze(0)
        h = torch.tanh(self.sap_linear(x))
        w = F.softmax(torch.matmul(h, self.attention).squeeze(dim=2), dim=1).view(x.size(0), x.size(1), 1)
        mu = torch.sum(x * w, dim=1)
        rh = torch.sqrt((torch.sum((x ** 2) * w, dim=1) - mu ** 2).clamp(min=1e-5))
        pooled.append(torch.cat((mu, rh), 1).squeeze(0))
    return torch.stack(pooled)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=36)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--autocast", action="store_true")
    ap.add_argument("--layerdrop", type=float, default=0.1)
    ap.add_argument("--keep-grads", action="store_true", help="fused optimizer: zero gradients in place (stable addresses)")
    ap.add_argument("--optimizer", choices=["fused", "torch"], default="fused")
    ap.add_argument("--stock-pool", action="store_true")
    args = ap.parse_args()

    rank, world, local_rank = init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    L = int(args.seconds * 16000)
    enc = WavLMEncoder(wavlm_large_config(num_hidden_layers=args.layers, layerdrop=args.layerdrop))
    model = EmotionClassifier(enc, hidden_dim=1024, dropout=0.3, num_emotions=8).to(dev)
    model.unfreeze_encoder_gradually(list(range(args.layers)))  # last fine-tuning epoch: every layer index
    if args.stock_pool:
        model.pooling.forward = stock_pool_forward.__get__(model.pooling)
    n_params = model.get_trainable_params()
    ddp = wrap_data_parallel(model, dev)
    if args.optimizer == "fused":
        opt = FusedAdamWEma(model.parameters(), lr=5e-6, weight_decay=1e-4, max_grad_norm=1.0)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=5e-6, weight_decay=1e-4)
    if args.optimizer == "fused":
        opt.keep_grads = args.keep_grads
    clean, noise, snr_idx, table = synthetic.waveforms(args.batch, L, seed=1234 + rank)
    raw = {"clean_wave": torch.from_numpy(clean)[:, None].pin_memory(), "noise_wave": torch.from_numpy(noise)[:, None].pin_memory(),
           "snr_idx": torch.from_numpy(snr_idx), "snr": torch.tensor([table[i] for i in snr_idx])}
    mixer = GpuBatchMixer([2, 5, 10, 15, 20], dev, peak_norm=False)
    g = torch.Generator().manual_seed(7 + rank)
    lens = torch.randint(L // 3, L + 1, (args.batch,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).float().to(dev)
    labels = (torch.rand(args.batch, 3, generator=g) * 6 + 1).to(dev)

    def step():
        batch = mixer(raw)
        x = batch["noisy_input_values"].squeeze(1)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.autocast):
            return emotion_dim_step(ddp, x, labels, opt, mask)[0]

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

    def ev(fn, n=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    h = torch.randn(args.batch, (L - 400) // 320 + 1, 1024, device=dev, requires_grad=True)

    def pool_fb():
        model.pooling(h, mask).sum().backward()
    t_pool = ev(pool_fb)
    if rank == 0:
        print(json.dumps({
            "workload": "configs[3]: emotion-dimension fine-tune step, WavLM-large shapes (random init), data-parallel",
            "n_gpus": world, "batch_per_gpu": args.batch, "seconds": args.seconds, "layers": args.layers,
            "autocast_bf16": args.autocast, "layerdrop": args.layerdrop, "optimizer": args.optimizer, "stock_pool": args.stock_pool,
            "trainable_params": n_params, "loss": float(loss), "ms_per_step": ms,
            "optimizer_table_builds": getattr(opt, "table_builds", None), "keep_grads": args.keep_grads,
            "utterance_seconds_per_s": world * args.batch * args.seconds / (ms * 1e-3),
            "hot_path_ms": {"attentive_pooling_fwd_bwd": t_pool},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
