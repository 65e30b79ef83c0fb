"""Device time of the mix-kernel variants (CUDA-graph timed, inputs rotated beyond L2).

usage: python scripts/bench_mix.py [--all]   (--all: cluster / carveout sweep and the generic fallback kernel)
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
for L, B, nsets in ((64000, 64, 6), (64000, 512, 2), (64000, 2048, 1), (80000, 36, 8), (80000, 512, 2), (192000, 256, 2)):
    clean, noise, snr_idx, table = synthetic.waveforms(min(B, 64), L, seed=1)
    tab = [float(v) for v in table]
    rep = max(1, B // clean.shape[0])
    sets = [(torch.from_numpy(clean).to(dev).repeat(rep, 1).contiguous() + 0.0 * i,
             torch.from_numpy(noise).to(dev).repeat(rep, 1).contiguous()) for i in range(nsets)]
    s = torch.from_numpy(snr_idx).to(dev).repeat(rep)
    Bt = sets[0][0].shape[0]
    combos = [(4, 0, -1), (5, 0, -1)]
    if "--all" in sys.argv:  # cluster / carveout sweep of the resident kernel and the generic fallback
        combos += [(4, cs, -1) for cs in (3, 4, 5, 8)] + [(4, 0, 100), (4, 0, 70), (0, 0, -1)]
    for variant, cs, carve in combos:
        ops.set_mix_variant(variant)
        ops.set_mix_cluster(cs)
        ops.set_mix_carveout(carve)
        n = 12
        for i in range(3):
            ops.mix_normalize(sets[i % nsets][0], sets[i % nsets][1], s, tab, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n):
                ops.mix_normalize(sets[i % nsets][0], sets[i % nsets][1], s, tab, True)
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        gbs = 16.0 * Bt * L / (best * 1e-3) / 1e9
        print(f"L={L:6d} B={Bt:5d} variant={variant} cs={cs} carveout={carve:3d}: {best*1e3:8.1f} us  {gbs:7.0f} GB/s  ({gbs/6555.2:.3f} of HBM peak)", flush=True)
    ops.set_mix_carveout(-1)
    ops.set_mix_cluster(0)
    del sets
ops.set_mix_variant(4)
