"""Layer-0 kernel variants at 64 x 4 s (CUDA-graph timed): 0 SIMT, 1 tensor core with LayerNorm + GELU epilogue, 2 tensor
core with LayerNorm folded into the GEMM operands (GELU-only epilogue), 3 = 2 with 16 epilogue warps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
B, L = 64, 64000
layers = synthetic.frontend_weights("layer", seed=0)
w = torch.from_numpy(layers[0]["conv"]).to(dev)
g = torch.from_numpy(layers[0]["gamma"]).to(dev)
b = torch.from_numpy(layers[0]["beta"]).to(dev)
x = torch.randn(B, L, device=dev)
T, P = ops.frontend_geometry(L)
for variant in (1, 2, 3, 0):
    ops.set_layer0_variant(variant)
    fn = lambda: ops.conv_layer0(x, w, g, b, "layer")
    fn(); torch.cuda.synchronize()
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
    byt = 4.0 * B * L + 2.0 * 512 * B * T[0]
    print(f"layer0 variant {variant}: {best*1e3:.1f} us  {byt/(best*1e-3)/1e9:.0f} GB/s ({byt/(best*1e-3)/1e9/6555.2:.3f} of HBM peak)", flush=True)
ops.set_layer0_variant(ops.DEFAULT_LAYER0_VARIANT)
