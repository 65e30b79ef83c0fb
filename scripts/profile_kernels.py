"""Small driver for ncu: every hot-path kernel at the BASELINE shape (64 x 4 s): mix (also at B=512), conv frontend
forward, one training forward + native backward, the fused optimizer tail (clip + AdamW + EMA, 80 M parameters) and
the batched attentive statistics pooling forward / backward (36 x 249 x 1024)."""
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
    yt, tape = ops.conv_frontend_train(c, w, g, b, "layer", packed=packed)
    grads = ops.conv_frontend_backward(c, w, g, b, tape, gy, "layer", dgrad_packs=dpacks)
from nrse_b200.train import FusedAdamWEma
from nrse_b200.models import AttentiveStatisticsPooling
prm = [torch.nn.Parameter(torch.randn(n, device=dev) * 0.02) for n in [1 << 22] * 19 + [1000, 12, 4097]]
twin = [torch.randn_like(p) for p in prm[:-1]]
for p in prm:
    p.grad = torch.randn_like(p) * 1e-3
opt = FusedAdamWEma(prm, lr=1e-5, weight_decay=1e-5, max_grad_norm=1.0, ema_pairs=zip(prm[:-1], twin), ema_decay=0.996)
pool = AttentiveStatisticsPooling(1024).to(dev)
hx = torch.randn(36, 249, 1024, device=dev, requires_grad=True)
mask = torch.ones(36, 80000, device=dev)
mask[1::2, 50000:] = 0
for it in range(2):
    opt.step()
    pool(hx, mask).sum().backward()
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(grads[0][3].abs().mean()), float(opt.last_grad_norm))
