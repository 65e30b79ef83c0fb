"""Parity of the conv feature encoder BACKWARD kernels with torch autograd on the fp32 oracle graph.
Gradients flow through bf16 activations / bf16 dZ, so the tolerance is the bf16 one of the north star (1e-2 relative,
norm-relative per tensor); the per-kernel tests are tighter where the operands are identical."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.utils import synthetic

pytestmark = pytest.mark.gpu


def test_ln_gelu_bwd_vs_autograd(dev):
    torch.manual_seed(0)
    P, T, B = 40, 37, 3
    rows = B * P
    z = (torch.randn(rows, 512) * 1.5 + 0.3).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(512)).requires_grad_(True)
    beta = (0.1 * torch.randn(512)).requires_grad_(True)
    out = F.gelu(F.layer_norm(z, (512,), gamma, beta, 1e-5))
    dout = torch.randn(rows, 512)
    valid = (torch.arange(rows) % P) < T
    dout = dout * valid[:, None]
    out.backward(dout)
    mean = z.detach().mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(z.detach().var(1, unbiased=False, keepdim=True) + 1e-5)
    xhat = ((z.detach() - mean) * rstd).bfloat16()
    for dtype in (torch.float32, torch.bfloat16):
        dz, dg, db = ops.ln_gelu_bwd(dout.to(dtype).to(dev), xhat.to(dev), rstd.flatten().to(dev), gamma.detach().to(dev),
                                     beta.detach().to(dev), P, T)
        tol = 1e-2 if dtype == torch.float32 else 2e-2
        assert rel_err(dz.float().cpu().numpy(), z.grad.numpy()) < tol
        assert rel_err(dg.cpu().numpy(), gamma.grad.numpy()) < tol
        assert rel_err(db.cpu().numpy(), beta.grad.numpy()) < tol
        assert not dz.float().cpu()[~valid].any()          # pitch padding carries zero gradient


@pytest.mark.parametrize("k,rows_out", [(3, 128), (2, 192), (3, 1000), (2, 64 * 131 + 7)])
def test_wgrad_and_dgrad_vs_autograd(dev, k, rows_out):
    rs = np.random.RandomState(k * 7 + rows_out)
    act = torch.from_numpy(rs.standard_normal((2 * rows_out, 512)).astype(np.float32)).bfloat16()
    w = torch.from_numpy((rs.standard_normal((512, 512, k)) * np.sqrt(2.0 / (512 * k))).astype(np.float32))
    dz = torch.from_numpy(rs.standard_normal((rows_out, 512)).astype(np.float32)).bfloat16()
    # torch reference on the same bf16-rounded operands (input zero-extended like the TMA out-of-bounds fill)
    a = torch.cat([act.float(), torch.zeros(2, 512)], 0).requires_grad_(True)
    wq = w.bfloat16().float().requires_grad_(True)
    h = F.conv1d(a.t()[None], wq, stride=2)[0].t()[:rows_out]
    h.backward(dz.float())
    dw_ref = wq.grad                                   # [512, 512, k]
    dx_ref = a.grad[:2 * rows_out]
    dw = ops.conv_layer_wgrad(dz.to(dev), act.to(dev), k)
    got_dw = dw.view(512, k, 512).permute(0, 2, 1).cpu()
    assert rel_err(got_dw.numpy(), dw_ref.numpy()) < 1e-3
    even, odd = ops.pack_conv_weight_dgrad(w.to(dev))
    dx = ops.conv_layer_dgrad(dz.to(dev), even, odd, k)
    assert rel_err(dx.float().cpu().numpy(), dx_ref.numpy()) < 6e-3   # bf16 output rounding


def test_layer0_wgrad_vs_autograd(dev):
    rs = np.random.RandomState(3)
    B, L = 3, 4000
    T, P = ops.frontend_geometry(L)
    x = torch.from_numpy(rs.standard_normal((B, L)).astype(np.float32))
    dz = torch.from_numpy(rs.standard_normal((B, P[0], 512)).astype(np.float32)).bfloat16()
    dz[:, T[0]:] = 0
    w = torch.zeros(512, 1, 10, requires_grad=True)
    y = F.conv1d(x[:, None], w, stride=5)              # [B, 512, T0]
    y.backward(dz[:, :T[0]].float().transpose(1, 2))
    got = ops.conv_layer0_wgrad(x.to(dev), dz.view(-1, 512).to(dev), T[0], P[0])
    assert rel_err(got.cpu().numpy(), w.grad[:, 0].numpy()) < 1e-4


@pytest.mark.parametrize("B,L", [(2, 4000), (3, 16000)])
def test_full_backward_vs_autograd(dev, B, L):
    layers = synthetic.frontend_weights("layer", seed=9)
    x = synthetic.waveforms(B, L, seed=5)[0]
    x = ((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype(np.float32)
    params = [{k: (torch.from_numpy(v).requires_grad_(True) if v is not None else None) for k, v in l.items()} for l in layers]
    y_ref = oracle.conv_frontend(torch.from_numpy(x), params, "layer")          # [B, 512, T]
    gy = torch.from_numpy(np.random.RandomState(1).standard_normal(tuple(y_ref.shape)).astype(np.float32))
    y_ref.backward(gy)

    w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
    g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
    b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
    xd = torch.from_numpy(x).to(dev)
    y, tape = ops.conv_frontend_train(xd, w, g, b)
    assert rel_err(y.transpose(1, 2).cpu().numpy(), y_ref.detach().numpy()) < 1e-2
    # the tape-writing forward runs layer 0 with the un-folded epilogue (it has to save the pre-affine activations): with
    # the same layer-0 kernel selected for inference the features are bit-identical, with the default (LayerNorm folded
    # into the GEMM operands) they agree to bf16 rounding of the layer-0 output
    ops.set_layer0_variant(1)
    ops.set_frontend_variant(2)   # the tape-writing forward always runs the 1-SM kernels
    assert torch.equal(y, ops.conv_frontend(xd, w, g, b, "layer"))
    ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)
    ops.set_layer0_variant(ops.DEFAULT_LAYER0_VARIANT)
    y_plain = ops.conv_frontend(xd, w, g, b, "layer")
    # two bf16 pipelines that differ in the rounding of layer 0, six layers later: each is within 1e-2 of the fp32 oracle
    assert rel_err(y_plain.transpose(1, 2).cpu().numpy(), y_ref.detach().numpy()) < 1e-2
    assert rel_err(y_plain.cpu().numpy(), y.cpu().numpy()) < 8e-3
    dw, dg, db = ops.conv_frontend_backward(xd, w, g, b, tape, gy.to(dev).transpose(1, 2))
    for i in range(7):
        e_w = rel_err(dw[i].cpu().numpy(), params[i]["conv"].grad.numpy())
        e_g = rel_err(dg[i].cpu().numpy(), params[i]["gamma"].grad.numpy())
        e_b = rel_err(db[i].cpu().numpy(), params[i]["beta"].grad.numpy())
        assert dw[i].shape == params[i]["conv"].shape
        # bf16 activations + bf16 gradient buffers across up to 7 layers: measured ~1e-2; bound 3e-2
        assert max(e_w, e_g, e_b) < 3e-2, (i, e_w, e_g, e_b)
