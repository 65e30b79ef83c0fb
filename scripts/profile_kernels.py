"""Small driver for ncu: every hot-path kernel at the BASELINE shape (64 x 4 s): mix (also at B=512), conv frontend
forward, then one training forward + native backward."""
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
dpacks = [ops.pack_conv_weight_dgrad(x) for x in w[1:]]
cb, nb, sb = c_d.repeat(8, 1).contiguous(), n_d.repeat(8, 1).contiguous(), s_d.repeat(8)
T, P = ops.frontend_geometry(L)
gy = torch.randn(B, T[6], 512, device=dev)
for it in range(2):
    c, n, st = ops.mix_normalize(c_d, n_d, s_d, tab, True)
    ops.mix_normalize(cb, nb, sb, tab, True)
    y = ops.conv_frontend(c, w, g, b, "layer", out_dtype=torch.bfloat16, packed=packed)
    yt, tape = ops.conv_frontend_train(c, w, g, b, packed=packed)
    grads = ops.conv_frontend_backward(c, w, g, b, tape, gy, dgrad_packs=dpacks)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(grads[0][3].abs().mean()))
