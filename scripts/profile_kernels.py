"""Small driver for ncu: a few launches of each hot-path kernel at the BASELINE shape (and mix at B=512)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
B, L = 64, 64000
clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=1234)
c_d, n_d, s_d = torch.from_numpy(clean).to(dev), torch.from_numpy(noise).to(dev), torch.from_numpy(snr_idx).to(dev)
tab = [float(v) for v in table]
layers = synthetic.frontend_weights("layer", seed=0)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
packed = [ops.pack_conv_weight(x) for x in w[1:]]
cb, nb, sb = c_d.repeat(8, 1).contiguous(), n_d.repeat(8, 1).contiguous(), s_d.repeat(8)
for it in range(3):
    c, n, st = ops.mix_normalize(c_d, n_d, s_d, tab, True)
    ops.mix_normalize(cb, nb, sb, tab, True)
    y = ops.conv_frontend(c, w, g, b, "layer", out_dtype=torch.bfloat16, packed=packed)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
