"""Graph-timed device time of layer 0 and the six GEMM layers at 64 x 4 s, for the GEMM kernel variants: 2 = 1-SM UMMA,
CTA pair splits the channels (default); 3 = 2-SM UMMA (cta_group::2), CTA pair splits the frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops, _lib
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
B, L = 64, 64000
layers = synthetic.frontend_weights("layer", seed=0)
x = synthetic.waveforms(B, L, seed=1)[0]
x = torch.from_numpy(((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype("float32")).to(dev)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
packed = [ops.pack_conv_weight(t) for t in w[1:]]
T, P = ops.frontend_geometry(L)
K = (10, 3, 3, 3, 3, 2, 2)

def gtime(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

lib = _lib.load()
for pf in (2, 3, 2, 3):
    ops.set_frontend_variant(pf)
    act = ops.conv_layer0(x, w[0], g[0], b[0], "layer").view(B * P[0], 512)
    t0 = gtime(lambda: ops.conv_layer0(x, w[0], g[0], b[0], "layer"))
    out = [f"l0={t0*1e3:.0f}us"]
    tot, fl = 0.0, 0.0
    for i in range(1, 7):
        inp = act
        t = gtime(lambda: ops.conv_layer(inp, packed[i - 1], K[i], g[i], b[i]))
        act = ops.conv_layer(inp, packed[i - 1], K[i], g[i], b[i])
        f = 2.0 * B * T[i] * 512 * 512 * K[i]
        tot += t; fl += f
        out.append(f"l{i}={t*1e3:.0f}us/{f/(t*1e-3)/1e12:.0f}TF")
    print(f"variant={pf}: " + " ".join(out) + f"  | gemm total {tot*1e3:.0f}us {fl/(tot*1e-3)/1e12:.0f} TF")
ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)
