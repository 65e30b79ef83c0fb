"""Experiment: run layers 0-2 of the conv frontend group by group (G utterances at a time) through small ring buffers
that stay resident in L2, then layers 3-6 over the whole batch -- against the layer-by-layer whole-batch schedule.
Motivation: the SM clock drops from ~1.8 to ~1.5 GHz while a layer's output streams to DRAM; with
ring buffers the 839 MB + 419 MB outputs of layers 0-1 are overwritten in L2 before they are ever written back."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops, _lib
from nrse_b200.ops import _ptr, _stream
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
B, L = 64, 64000
layers = synthetic.frontend_weights("layer", seed=0)
x = torch.randn(B, L, device=dev)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
packed = [ops.pack_conv_weight(t) for t in w[1:]]
T, P = ops.frontend_geometry(L)
K = (10, 3, 3, 3, 3, 2, 2)
lib = _lib.load()
w0 = w[0].reshape(512, 10).contiguous()
full = [torch.empty(B * P[i], 512, dtype=torch.bfloat16, device=dev) for i in range(7)]

def l0(xs, out, nb):
    _lib.check(lib.nrse_conv_layer0_fwd(_ptr(xs), _ptr(w0), _ptr(g[0]), _ptr(b[0]), 0, _ptr(out), None, nb, L, T[0], P[0], _stream()))

def li(i, inp, out, rows_out):
    _lib.check(lib.nrse_conv_layer_fwd(_ptr(inp), 2 * rows_out, _ptr(packed[i - 1]), K[i], 2, _ptr(g[i]), _ptr(b[i]), _ptr(out), 1, rows_out, _stream()))

def whole(v12):
    l0(x, full[0], B)
    for i in range(1, 7):
        ops.set_frontend_variant(v12 if i <= 2 else 2)
        li(i, full[i - 1], full[i], B * P[i])

def grouped(G, depth, v12, rings):
    for s in range(0, B, G):
        nb = min(G, B - s)
        l0(x[s:s + nb], rings[0], nb)
        prev = rings[0]
        for i in range(1, depth + 1):
            ops.set_frontend_variant(v12)
            out = full[i][s * P[i]:(s + nb) * P[i]] if i == depth else rings[i]
            li(i, prev, out, nb * P[i])
            prev = out
    for i in range(depth + 1, 7):
        ops.set_frontend_variant(2)
        li(i, full[i - 1], full[i], B * P[i])

def gtime(fn, n=3):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best

whole(3); torch.cuda.synchronize()
ref = full[6].clone()
for v in (2, 3):
    print(f"whole batch, layers 1-2 on variant {v}: {gtime(lambda: whole(v))*1e3:.0f} us", flush=True)
for depth in (1, 2, 3):
    for G in (2, 4, 5, 8, 16):
        rings = [torch.empty(G * P[i], 512, dtype=torch.bfloat16, device=dev) for i in range(depth)]
        mb = sum(r.numel() * 2 for r in rings) / 1e6
        for v in (2, 3):
            t = gtime(lambda: grouped(G, depth, v, rings))
            same = torch.equal(full[6], ref)
            print(f"groups of {G:2d}, layers 0-{depth} grouped (rings {mb:.0f} MB), variant {v}: {t*1e3:.0f} us  same result: {same}", flush=True)
ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)
