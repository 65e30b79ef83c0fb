"""CTAs-per-row sweep of the resident mix kernel over utterance lengths (feeds the dispatch rule in mix.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")


def timeit(sets, s, tab, n=10):
    ns = len(sets)
    for i in range(2):
        ops.mix_normalize(sets[i % ns][0], sets[i % ns][1], s, tab, True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            ops.mix_normalize(sets[i % ns][0], sets[i % ns][1], s, tab, True)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best


for L in (16000, 32000, 48000, 64000, 80000, 96000, 128000, 160000, 192000):
    B = max(64, (512 * 64000 // L) // 64 * 64)
    clean, noise, snr_idx, table = synthetic.waveforms(64, L, seed=1)
    tab = [float(v) for v in table]
    rep = B // 64
    sets = [(torch.from_numpy(clean).to(dev).repeat(rep, 1).contiguous() + 0.0 * i,
             torch.from_numpy(noise).to(dev).repeat(rep, 1).contiguous()) for i in range(2)]
    s = torch.from_numpy(snr_idx).to(dev).repeat(rep)
    nvec = L // 4
    out = []
    ops.set_mix_variant(3); ops.set_mix_cluster(0)
    out.append(("stream", 0, timeit(sets, s, tab)))
    for variant, cap in ((4, 2048 + 3456), (5, 4096 + 6144)):
        ops.set_mix_variant(variant)
        for cs in range(1, 9):
            if -(-nvec // cs) > cap:
                continue
            ops.set_mix_cluster(cs)
            out.append(("r512" if variant == 4 else "r1024", cs, timeit(sets, s, tab)))
    ops.set_mix_cluster(0)
    bytes_ = 16.0 * B * L
    print(f"L={L} B={B}: " + "  ".join(f"{n}/{cs}:{bytes_/(t*1e-3)/1e9/6555.2:.3f}" for n, cs, t in out), flush=True)
    del sets
ops.set_mix_variant(4)
