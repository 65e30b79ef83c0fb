#!/usr/bin/env python
"""BASELINE.json configs[4]: embedding-extraction sweep of evaluate_byol.py -- clean and noisy views at SNR 4 / 8 dB for
utterance lengths 2-12 s -- through the hot path (GPU mix + conv feature encoder forward on both views, no grad).
The reference has no length masking: every run uses one fixed L (ref:evaluate_byol.py:12-66), so the
sweep is a loop over L.  Prints one JSON line per length: device time, utterance-seconds/s, cosine of pooled conv
features between the two views per SNR (a sanity number, random weights)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nrse_b200 import ops  # noqa: E402
from nrse_b200.train import init_distributed  # noqa: E402
from nrse_b200.utils import synthetic  # noqa: E402


def main():
    """Single process, or one process per GPU under torchrun: the utterances are sharded over the ranks (64 per rank,
    different seeds), the per-SNR similarity sums meet in ONE all-reduce per length."""
    rank, world, local_rank = init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    layers = synthetic.frontend_weights("layer", seed=0)
    w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
    g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
    b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
    packed = [ops.pack_conv_weight(t) for t in w[1:]]
    snr_table = [4.0, 8.0]
    B = 64
    for seconds in (2, 4, 6, 8, 10, 12):
        L = seconds * 16000
        clean, noise, snr_idx, _ = synthetic.waveforms(B, L, seed=seconds + 100 * rank, snr_range=(4, 8))
        c_d, n_d, s_d = (torch.from_numpy(a).to(dev) for a in (clean, noise, snr_idx))

        def run():
            c, n, st = ops.mix_normalize(c_d, n_d, s_d, snr_table, True)
            yc = ops.conv_frontend(c, w, g, b, "layer", out_dtype=torch.bfloat16, packed=packed)
            yn = ops.conv_frontend(n, w, g, b, "layer", out_dtype=torch.bfloat16, packed=packed)
            return yc, yn, st
        for _ in range(3):
            yc, yn, st = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_it = 10
        e0.record()
        for _ in range(n_it):
            yc, yn, st = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_it
        sims = ops.cosine_rows_plain(yc.float().mean(1), yn.float().mean(1))
        acc = torch.stack([torch.stack([sims[s_d == i].sum(), (s_d == i).sum().float()]) for i in range(len(snr_table))])
        stats = torch.tensor([ms, float((st != 0).sum())], device=dev)
        if world > 1:
            dist.all_reduce(acc)                                  # per-SNR (sum, count) over all ranks
            t = stats[:1].clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)              # the slowest rank defines the time
            dist.all_reduce(stats[1:])
            stats[0] = t[0]
        ms = float(stats[0])
        per_snr = {str(int(v)): float(acc[i, 0] / acc[i, 1].clamp_min(1)) for i, v in enumerate(snr_table)}
        if rank == 0:
            print(json.dumps({"seconds": seconds, "n_gpus": world, "batch_per_gpu": B, "frames": int(yc.shape[1]), "ms": ms,
                              "utterance_seconds_per_s": world * B * seconds / (ms * 1e-3), "rejected_rows": int(stats[1]),
                              "pooled_conv_feature_cosine_by_snr": per_snr}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
