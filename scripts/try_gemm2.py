"""Bring-up check of the 2-SM UMMA forward kernel (frontend variant 3) against the default kernel (variant 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
layers = synthetic.frontend_weights("layer", seed=0)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
packed = [ops.pack_conv_weight(t) for t in w[1:]]
K = (10, 3, 3, 3, 3, 2, 2)
for rows_out, k in ((128, 3), (256, 3), (384, 2), (1000, 3), (128 * 149 + 5, 3)):
    act = (torch.randn(2 * rows_out, 512, device=dev) * 0.5).bfloat16()
    li = 1 if k == 3 else 5
    ops.set_frontend_variant(2)
    ref = ops.conv_layer(act, packed[li - 1], k, g[li], b[li]).float()
    torch.cuda.synchronize()
    ops.set_frontend_variant(3)
    got = ops.conv_layer(act, packed[li - 1], k, g[li], b[li]).float()
    torch.cuda.synchronize()
    d = (got - ref).abs().max().item()
    print(f"rows_out={rows_out} k={k}: max|diff| = {d:.3e}  (ref absmax {ref.abs().max().item():.3f})  equal={torch.equal(got, ref)}", flush=True)
ops.set_frontend_variant(2)
print("done")
