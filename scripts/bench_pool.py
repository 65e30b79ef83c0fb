"""Attentive statistics pooling at the emotion fine-tune shape (36 x 249 x 1024) and at 64 x 199 x 1024: device time of
the fused forward / backward kernels (CUDA-graph timed, sap_linear GEMM excluded) and of the reference-style loop."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from nrse_b200 import ops

dev = torch.device("cuda:0")
HBM = 6555.2


def graph_ms(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = []
for B, T, D in ((36, 249, 1024), (64, 199, 1024), (256, 249, 1024)):
    x = torch.randn(B, T, D, device=dev)
    hl = torch.randn(B, T, D, device=dev)
    att = torch.randn(D, device=dev) * 0.03
    lens = torch.randint(T // 2, T + 1, (B,), device=dev, dtype=torch.int32)
    lens[0] = T
    frames = int(lens.sum())
    o, w = ops._asp_pool_fwd(x, hl, att, lens)
    go = torch.randn_like(o)
    t_f = graph_ms(lambda: ops._asp_pool_fwd(x, hl, att, lens))
    t_b = graph_ms(lambda: ops._asp_pool_bwd(x, hl, att, lens, o, w, go))
    # algorithmic bytes over the VALID frames: fwd reads hl and x once; bwd reads x twice and hl once, writes dx and dhl
    # for all B*T frames (padded frames are zero-filled)
    bf = 8.0 * frames * D
    bb = 12.0 * frames * D + 8.0 * B * T * D
    out.append({"shape": [B, T, D], "valid_frames": frames, "fwd_ms": t_f, "bwd_ms": t_b,
                "fwd_GBs": bf / (t_f * 1e-3) / 1e9, "fwd_frac": bf / (t_f * 1e-3) / 1e9 / HBM,
                "bwd_GBs": bb / (t_b * 1e-3) / 1e9, "bwd_frac": bb / (t_b * 1e-3) / 1e9 / HBM})
print(json.dumps(out))
