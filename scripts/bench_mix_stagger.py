"""Start-stagger sweep of the single-wave resident mix launch (B x 4 s rows resident at once)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
for L, B in ((64000, 64), (64000, 32), (64000, 16), (80000, 36), (32000, 64), (64000, 74), (64000, 128)):
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=1)
    tab = [float(v) for v in table]
    nsets = 6
    sets = [(torch.from_numpy(clean).to(dev) + 0.0 * i, torch.from_numpy(noise).to(dev).clone()) for i in range(nsets)]
    s = torch.from_numpy(snr_idx).to(dev)
    res = []
    for variant in (4, 5):
        ops.set_mix_variant(variant)
        for groups in (0, 2, 3, 4, 6, 8):
            ops.set_mix_stagger(groups)
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
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / n)
            res.append(f"v{variant}/g{groups}:{best*1e3:.1f}us({16.0*B*L/(best*1e-3)/1e9/6555.2:.3f})")
    print(f"L={L} B={B}: " + "  ".join(res), flush=True)
ops.set_mix_stagger(-1); ops.set_mix_variant(4)
