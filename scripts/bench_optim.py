"""Optimizer tail of the BYOL step at WavLM-large size: FusedAdamWEma (2 launches) vs the reference's sequence in
stock torch on the same GPU (clip_grad_norm_ + torch.optim.AdamW [foreach, the CUDA default] + per-tensor EMA loop)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200.train import FusedAdamWEma

dev = torch.device("cuda:0")
# encoder + projector (have EMA twins): 317,556,416; predictor (no twin): 8,401,920  => 325,958,336 trainable
twin_sizes = [8] * 24 + [16] * 24 + [128] + [512] * 40 + [1024] * 227 + [4096] * 24 + [5120] * 2 + [524288] * 3 + \
             [786432] * 4 + [1048576] * 98 + [4194304] * 48 + [8388608]
pred_sizes = [1024 * 2048, 2048, 2048, 2048, 2048 * 2048, 2048, 2048, 2048, 2048 * 1024, 1024]
HBM = 6555.2


def time_ms(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def graph_ms(fn, n=10):
    """Device time: n calls captured in one CUDA graph (no host gaps), replayed between events."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def make():
    torch.manual_seed(0)
    P = [torch.nn.Parameter(torch.randn(n, device=dev) * 0.02) for n in twin_sizes + pred_sizes]
    T = [torch.randn(n, device=dev) * 0.02 for n in twin_sizes]
    for p in P:
        p.grad = torch.randn_like(p) * 1e-3
    return P, T


out = {}
P, T = make()
opt = FusedAdamWEma(P, lr=1e-5, weight_decay=1e-5, max_grad_norm=1.0, ema_pairs=zip(P[:len(T)], T), ema_decay=0.996)
opt.step()  # builds state + table
torch.cuda.synchronize()
ms_host = time_ms(opt.step)
ms = graph_ms(opt.step)
n_upd, n_ema = sum(p.numel() for p in P), sum(t.numel() for t in T)
alg = 4.0 * n_upd + 28.0 * n_upd + 8.0 * n_ema   # norm read + (p,g,m,v read; p,m,v write) + (t read, t write)
out["fused"] = {"ms": ms, "ms_host_launched": ms_host, "params": n_upd, "ema_params": n_ema, "launches": 2, "algorithmic_bytes": alg,
                "achieved_GBs": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / HBM}
# the update kernel alone (no clip): 28 B + 8 B per parameter
opt.max_grad_norm = 0.0
ms2 = graph_ms(opt.step)
alg2 = 28.0 * n_upd + 8.0 * n_ema
out["fused_no_clip"] = {"ms": ms2, "launches": 1, "algorithmic_bytes": alg2, "achieved_GBs": alg2 / (ms2 * 1e-3) / 1e9,
                        "frac_of_hbm_peak": alg2 / (ms2 * 1e-3) / 1e9 / HBM}
del opt, P, T
torch.cuda.empty_cache()

P, T = make()
ref = torch.optim.AdamW(P, lr=1e-5, weight_decay=1e-5)   # ref:train_byol.py:146 (foreach on CUDA)


def stock():
    torch.nn.utils.clip_grad_norm_(P, 1.0)               # ref:train_byol.py:67
    ref.step()                                           # :70
    for i in range(len(T)):                              # ref:src/models/byol.py:64-73
        T[i] = 0.996 * T[i] + (1 - 0.996) * P[i].data


out["stock_torch_same_gpu"] = {"ms": time_ms(stock, n=5, warm=2)}
del ref
torch.cuda.empty_cache()
fz = torch.optim.AdamW(P, lr=1e-5, weight_decay=1e-5, fused=True)


def stock_fused():
    torch.nn.utils.clip_grad_norm_(P, 1.0)
    fz.step()
    for i in range(len(T)):
        T[i] = 0.996 * T[i] + (1 - 0.996) * P[i].data


out["stock_torch_fused_adamw_same_gpu"] = {"ms": time_ms(stock_fused, n=5, warm=2)}
out["speedup_vs_stock"] = out["stock_torch_same_gpu"]["ms"] / out["fused"]["ms"]
print(json.dumps(out))
