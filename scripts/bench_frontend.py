"""Whole conv-frontend forward (layer 0 + six GEMM layers, 64 x 4 s, bf16 out), CUDA-graph timed, for the tile-order
knob: 0 = every layer walks its tiles first-to-last, 1 = consecutive layers in opposite directions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
B, L = 64, 64000
layers = synthetic.frontend_weights("layer", seed=0)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
packed = [ops.pack_conv_weight(t) for t in w[1:]]
x = torch.randn(B, L, device=dev)
ref = None
for order in (0, 1, 0, 1):
    ops.set_tile_order(order)
    fn = lambda: ops.conv_frontend(x, w, g, b, "layer", out_dtype=torch.bfloat16, packed=packed)
    y = fn(); torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    same = torch.equal(y, ref)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            fn()
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"tile order {order}: frontend forward {best*1e3:.0f} us  ({B*4/(best*1e-3)/1e3:.1f} k utterance-s/s per view)  bit-identical to order 0: {same}", flush=True)
ops.set_tile_order(1)
