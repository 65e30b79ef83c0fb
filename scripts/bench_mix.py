"""Device time of every mix-kernel variant at B=64 and B=512 (CUDA-graph timed, inputs rotated beyond L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
L = 64000
clean, noise, snr_idx, table = synthetic.waveforms(64, L, seed=1)
tab = [float(v) for v in table]
for B, nsets in ((64, 6), (512, 2), (2048, 1)):
    rep = B // 64
    sets = [(torch.from_numpy(clean).to(dev).repeat(rep, 1).contiguous() + 0.0 * i,
             torch.from_numpy(noise).to(dev).repeat(rep, 1).contiguous()) for i in range(nsets)]
    s = torch.from_numpy(snr_idx).to(dev).repeat(rep)
    for variant in (3, 2, 1, 0):
        ops.set_mix_variant(variant)
        n = 12
        for i in range(3):
            ops.mix_normalize(sets[i % nsets][0], sets[i % nsets][1], s, tab, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n):
                ops.mix_normalize(sets[i % nsets][0], sets[i % nsets][1], s, tab, True)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B:5d} variant={variant}: {ms*1e3:8.1f} us  {16.0*B*L/(ms*1e-3)/1e9:7.0f} GB/s  ({16.0*B*L/(ms*1e-3)/1e9/6555.2:.3f} of HBM peak)")
    del sets
