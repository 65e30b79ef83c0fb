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
    # fp32 gradient in the COMPACT layout [B, T, 512] (what autograd hands over: no zero-padded copy is ever made)
    compact = dout.view(B, P, 512)[:, :T].contiguous()
    dz, dg, db = ops.ln_gelu_bwd(compact.to(dev), xhat.to(dev), rstd.flatten().to(dev), gamma.detach().to(dev),
                                 beta.detach().to(dev), P, T, dout_pitch=T)
    assert rel_err(dz.float().cpu().numpy(), z.grad.numpy()) < 1e-2
    assert rel_err(dg.cpu().numpy(), gamma.grad.numpy()) < 1e-2 and rel_err(db.cpu().numpy(), beta.grad.numpy()) < 1e-2


def test_gelu_bwd_without_norm_vs_autograd(dev):
    """GroupNorm-mode layers 1-6 have no normalisation (hf:...modeling_wavlm.py:682-700): dZ = dOut gelu'(Z)."""
    torch.manual_seed(1)
    P, T, B = 24, 21, 4
    rows = B * P
    z = (torch.randn(rows, 512) * 1.5).bfloat16().float().requires_grad_(True)
    dout = torch.randn(rows, 512).bfloat16().float() * ((torch.arange(rows) % P) < T)[:, None]
    F.gelu(z).backward(dout)
    for dtype in (torch.float32, torch.bfloat16):
        dz, dg, db = ops.ln_gelu_bwd(dout.to(dtype).to(dev), z.detach().bfloat16().to(dev), None, None, None, P, T)
        assert dg is None and db is None
        assert rel_err(dz.float().cpu().numpy(), z.grad.numpy()) < 6e-3   # bf16 output rounding + the gelu' fit (2.6e-5)


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
    # the same accumulators written straight into the checkpoint layout [512, 512, k] (what the full backward uses)
    dw_ck = ops.conv_layer_wgrad(dz.to(dev), act.to(dev), k, ckpt_layout=True).cpu()
    assert dw_ck.shape == dw_ref.shape and rel_err(dw_ck.numpy(), dw_ref.numpy()) < 1e-3
    even, odd = ops.pack_conv_weight_dgrad(w.to(dev))
    dx = ops.conv_layer_dgrad(dz.to(dev), even, odd, k)
    assert rel_err(dx.float().cpu().numpy(), dx_ref.numpy()) < 6e-3   # bf16 output rounding


@pytest.mark.parametrize("k,B,P_prev,T_prev", [(3, 3, 128, 121), (2, 2, 256, 250), (3, 5, 640, 639), (2, 64, 24, 21)])
def test_dgrad_with_fused_ln_gelu_bwd_vs_autograd(dev, k, B, P_prev, T_prev):
    """One kernel: dX = dZ W (transposed stride-2 conv) -> LayerNorm + GELU backward of the layer below
    (hf:models/wavlm/modeling_wavlm.py:250-275 twice over, through autograd).  Checked against autograd on the same
    operands and against the two-kernel native path it replaces."""
    rs = np.random.RandomState(k * 11 + P_prev)
    rows_prev, rows_out = B * P_prev, B * P_prev // 2
    valid = torch.from_numpy((np.arange(rows_prev) % P_prev) < T_prev)
    zp = torch.from_numpy((rs.standard_normal((rows_prev, 512)) * 1.5 + 0.3).astype(np.float32)).requires_grad_(True)
    gamma = torch.from_numpy((1 + 0.1 * rs.standard_normal(512)).astype(np.float32)).requires_grad_(True)
    beta = torch.from_numpy((0.1 * rs.standard_normal(512)).astype(np.float32)).requires_grad_(True)
    w = torch.from_numpy((rs.standard_normal((512, 512, k)) * np.sqrt(2.0 / (512 * k))).astype(np.float32))
    dz = torch.from_numpy(rs.standard_normal((rows_out, 512)).astype(np.float32)).bfloat16()
    # the layer above sees zero gradient on its own padding frames (as the real backward guarantees)
    T_out = (T_prev - k) // 2 + 1
    dz = dz * torch.from_numpy((np.arange(rows_out) % (P_prev // 2)) < T_out)[:, None]
    act = F.gelu(F.layer_norm(zp, (512,), gamma, beta, 1e-5)) * valid[:, None]
    a = torch.cat([act, torch.zeros(2, 512)], 0)
    h = F.conv1d(a.t()[None], w.bfloat16().float(), stride=2)[0].t()[:rows_out]
    h.backward(dz.float())
    mean = zp.detach().mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(zp.detach().var(1, unbiased=False, keepdim=True) + 1e-5)
    xhat = ((zp.detach() - mean) * rstd).bfloat16().to(dev)
    rstd_d, g_d, b_d = rstd.flatten().to(dev), gamma.detach().to(dev), beta.detach().to(dev)
    even, odd = ops.pack_conv_weight_dgrad(w.to(dev))
    got, dg, db = ops.conv_layer_dgrad_lnbwd(dz.to(dev), even, odd, k, xhat, rstd_d, g_d, b_d, P_prev, T_prev)
    ref = zp.grad * valid[:, None]
    assert rel_err(got.float().cpu().numpy(), ref.numpy()) < 1e-2
    assert rel_err(dg.cpu().numpy(), gamma.grad.numpy()) < 1e-2
    assert rel_err(db.cpu().numpy(), beta.grad.numpy()) < 1e-2
    assert not got.float().cpu()[~valid].any()              # pitch padding carries zero gradient
    # against the two kernels it replaces (those round dX to bf16 in between; the fused form keeps it in fp32)
    dx = ops.conv_layer_dgrad(dz.to(dev), even, odd, k)
    two, dg2, db2 = ops.ln_gelu_bwd(dx, xhat, rstd_d, g_d, b_d, P_prev, T_prev)
    assert rel_err(got.float().cpu().numpy(), two.float().cpu().numpy()) < 8e-3
    assert rel_err(dg.cpu().numpy(), dg2.cpu().numpy()) < 8e-3 and rel_err(db.cpu().numpy(), db2.cpu().numpy()) < 8e-3
    # without the affine gradients (frozen LayerNorm): same dZ, nothing else written
    only, none_g, none_b = ops.conv_layer_dgrad_lnbwd(dz.to(dev), even, odd, k, xhat, rstd_d, g_d, b_d, P_prev, T_prev,
                                                      want_affine=False)
    assert none_g is None and none_b is None and torch.equal(only, got)


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


def _oracle_grads(x, layers, norm_mode, gy):
    params = [{k: (torch.from_numpy(v).requires_grad_(True) if v is not None else None) for k, v in l.items()} for l in layers]
    y_ref = oracle.conv_frontend(torch.from_numpy(x), params, norm_mode)          # [B, 512, T]
    y_ref.backward(gy)
    return y_ref.detach(), params


def _dev_params(layers, dev, n_norm):
    w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
    g = [torch.from_numpy(layers[i]["gamma"]).to(dev) for i in range(n_norm)]
    b = [torch.from_numpy(layers[i]["beta"]).to(dev) for i in range(n_norm)]
    return w, g, b


# Tolerance of the FULL backward against autograd on the fp32 oracle graph: 2e-2 norm-relative per tensor (it was 3e-2).
# The per-kernel tests on identical operands are held to 1e-3 (wgrad), 6e-3 (dgrad) and 1e-2 (norm + GELU backward); the
# full backward chains them: the gradient reaches layer i through (6 - i) data-gradient GEMMs with bf16 operands and bf16
# gradient buffers, on top of a forward whose bf16 activations / xhat are themselves within 5e-3 of the oracle.  Measured
# on the B200 (printed by the tests; profiles/README.md): 1.0-1.2e-2 for the weight gradients and up to 1.8e-2 for dgamma
# (a sum of products of two bf16-rounded tensors) at 8 x 64 000; it does not grow with depth, so one bound for all layers.
FULL_BWD_TOL = {i: 2.0e-2 for i in range(7)}


@pytest.fixture(params=[False, True], ids=["separate", "fused"])
def bwd_fusion(request):
    """Both forms of the LayerNorm-mode backward: LayerNorm / GELU backward as kernels of their own, or inside the
    data-gradient epilogue of the layer above."""
    ops.set_bwd_fusion(request.param)
    yield request.param
    ops.set_bwd_fusion(ops.DEFAULT_BWD_FUSION)


@pytest.mark.parametrize("B,L", [(2, 4000), (3, 16000)])
def test_full_backward_vs_autograd(dev, B, L, bwd_fusion):
    layers = synthetic.frontend_weights("layer", seed=9)
    x = synthetic.waveforms(B, L, seed=5)[0]
    x = ((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype(np.float32)
    T, _ = ops.frontend_geometry(L)
    gy = torch.from_numpy(np.random.RandomState(1).standard_normal((B, 512, T[6])).astype(np.float32))
    y_ref, params = _oracle_grads(x, layers, "layer", gy)
    w, g, b = _dev_params(layers, dev, 7)
    xd = torch.from_numpy(x).to(dev)
    y, tape = ops.conv_frontend_train(xd, w, g, b)
    assert rel_err(y.transpose(1, 2).cpu().numpy(), y_ref.numpy()) < 1e-2
    # the tape-writing forward runs layer 0 with the un-folded epilogue (it has to save the pre-affine activations) and
    # the same GEMM kernels as inference (2-SM for layers 1-3): with the un-folded layer-0 kernel selected for inference
    # the features are bit-identical, with the default (LayerNorm folded into the GEMM operands) they agree to bf16
    # rounding of the layer-0 output
    ops.set_layer0_variant(1)
    assert torch.equal(y, ops.conv_frontend(xd, w, g, b, "layer"))
    ops.set_layer0_variant(ops.DEFAULT_LAYER0_VARIANT)
    y_plain = ops.conv_frontend(xd, w, g, b, "layer")
    # two bf16 pipelines that differ in the rounding of layer 0, six layers later: each is within 1e-2 of the fp32 oracle
    assert rel_err(y_plain.transpose(1, 2).cpu().numpy(), y_ref.numpy()) < 1e-2
    assert rel_err(y_plain.cpu().numpy(), y.cpu().numpy()) < 8e-3
    dw, dg, db = ops.conv_frontend_backward(xd, w, g, b, tape, gy.to(dev).transpose(1, 2))
    for i in range(7):
        e_w = rel_err(dw[i].cpu().numpy(), params[i]["conv"].grad.numpy())
        e_g = rel_err(dg[i].cpu().numpy(), params[i]["gamma"].grad.numpy())
        e_b = rel_err(db[i].cpu().numpy(), params[i]["beta"].grad.numpy())
        print(f"full backward {B}x{L} layer {i}: dW {e_w:.2e} dgamma {e_g:.2e} dbeta {e_b:.2e}")
        assert dw[i].shape == params[i]["conv"].shape
        assert max(e_w, e_g, e_b) < FULL_BWD_TOL[i], (i, e_w, e_g, e_b)


@pytest.mark.parametrize("B,L", [(2, 4000), (3, 9000)])
def test_full_backward_group_mode_vs_autograd(dev, B, L):
    """GroupNorm mode (wavlm-base, hf:...modeling_wavlm.py:730-751 + :682-700): native training forward + backward --
    GELU-only backward for layers 1-6, one fused pass for layer 0 (GroupNorm-over-time backward + dW0)."""
    layers = synthetic.frontend_weights("group", seed=4)
    x = synthetic.waveforms(B, L, seed=6)[0]
    x = ((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype(np.float32)
    T, _ = ops.frontend_geometry(L)
    gy = torch.from_numpy(np.random.RandomState(2).standard_normal((B, 512, T[6])).astype(np.float32))
    y_ref, params = _oracle_grads(x, layers, "group", gy)
    w, g, b = _dev_params(layers, dev, 1)
    xd = torch.from_numpy(x).to(dev)
    y, tape = ops.conv_frontend_train(xd, w, g, b, "group")
    assert rel_err(y.transpose(1, 2).cpu().numpy(), y_ref.numpy()) < 1e-2
    assert torch.equal(y, ops.conv_frontend(xd, w, g, b, "group"))   # same kernels, the tape is extra stores
    dw, dg, db = ops.conv_frontend_backward(xd, w, g, b, tape, gy.to(dev).transpose(1, 2), "group")
    assert len(dg) == 1 and len(db) == 1
    for i in range(7):
        e = rel_err(dw[i].cpu().numpy(), params[i]["conv"].grad.numpy())
        print(f"group-mode backward {B}x{L} layer {i}: dW {e:.2e}")
        assert dw[i].shape == params[i]["conv"].shape and e < FULL_BWD_TOL[i], (i, e)
    e_g = rel_err(dg[0].cpu().numpy(), params[0]["gamma"].grad.numpy())
    e_b = rel_err(db[0].cpu().numpy(), params[0]["beta"].grad.numpy())
    assert max(e_g, e_b) < FULL_BWD_TOL[0], (e_g, e_b)


def test_backward_honours_needs_grad_mask(dev, bwd_fusion):
    """Partial unfreeze (ref:src/models/emotion.py:114-129): only what is asked for is computed; what IS computed equals
    the corresponding tensors of the full backward (same kernels, same order), the rest comes back as None."""
    B, L = 2, 6000
    layers = synthetic.frontend_weights("layer", seed=9)
    x = synthetic.waveforms(B, L, seed=8)[0]
    w, g, b = _dev_params(layers, dev, 7)
    xd = torch.from_numpy(x).to(dev)
    T, _ = ops.frontend_geometry(L)
    gy = torch.randn(B, T[6], 512, device=dev)
    _, tape = ops.conv_frontend_train(xd, w, g, b)
    full = ops.conv_frontend_backward(xd, w, g, b, tape, gy)
    need_w = [False, False, False, False, True, False, True]
    need_a = [False, False, False, False, False, True, True]
    dw, dg, db = ops.conv_frontend_backward(xd, w, g, b, tape, gy, need_w=need_w, need_affine=need_a)
    for i in range(7):
        assert (dw[i] is not None) == need_w[i] and (dg[i] is not None) == need_a[i] and (db[i] is not None) == need_a[i]
        if need_w[i]:  # split-K partial sums meet through fp32 atomics: equal up to summation order
            assert rel_err(dw[i].cpu().numpy(), full[0][i].cpu().numpy()) < 1e-5
        if need_a[i]:
            assert rel_err(dg[i].cpu().numpy(), full[1][i].cpu().numpy()) < 1e-5
            assert rel_err(db[i].cpu().numpy(), full[2][i].cpu().numpy()) < 1e-5
    nothing = ops.conv_frontend_backward(xd, w, g, b, tape, gy, need_w=[False] * 7, need_affine=[False] * 7)
    assert all(t is None for part in nothing for t in part)


def test_full_size_backward(dev, bwd_fusion):
    """BASELINE shape 64 x 64 000 (split-K weight gradients over M = 409 536 frames, fp32 atomics, 3.3 GB tape).
    (1) Frame independence: the gradients of the whole batch equal the SUM of the gradients of its eight 8-utterance
    chunks run separately (differences = fp32 summation order only).  (2) Chunk 0 against torch autograd on the fp32
    oracle graph at full utterance length, every tensor of every layer, at the tolerance of the small-shape test."""
    B, L, CH = 64, 64000, 8
    layers = synthetic.frontend_weights("layer", seed=0)
    x = synthetic.waveforms(B, L, seed=1234)[0]
    x = ((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype(np.float32)
    T, _ = ops.frontend_geometry(L)
    gy = torch.from_numpy(np.random.RandomState(11).standard_normal((B, T[6], 512)).astype(np.float32))
    w, g, b = _dev_params(layers, dev, 7)
    xd, gyd = torch.from_numpy(x).to(dev), gy.to(dev)
    _, tape = ops.conv_frontend_train(xd, w, g, b)
    full = ops.conv_frontend_backward(xd, w, g, b, tape, gyd)
    full = [[t.clone() for t in part] for part in full]
    del tape
    acc = None
    first = None
    for c0 in range(0, B, CH):
        _, tp = ops.conv_frontend_train(xd[c0:c0 + CH], w, g, b)
        part = ops.conv_frontend_backward(xd[c0:c0 + CH], w, g, b, tp, gyd[c0:c0 + CH])
        part = [[t.double() for t in p] for p in part]
        if first is None:
            first = [[t.clone() for t in p] for p in part]
        acc = part if acc is None else [[a + t for a, t in zip(pa, pt)] for pa, pt in zip(acc, part)]
        del tp
    for name, fa, aa in zip(("dW", "dgamma", "dbeta"), full, acc):
        for i in range(7):
            e = rel_err(fa[i].cpu().numpy(), aa[i].cpu().numpy())
            assert e < 2e-4, (name, i, e)
    _, params = _oracle_grads(x[:CH], layers, "layer", gy[:CH].transpose(1, 2).contiguous())
    for i in range(7):
        e_w = rel_err(first[0][i].cpu().numpy(), params[i]["conv"].grad.numpy())
        e_g = rel_err(first[1][i].cpu().numpy(), params[i]["gamma"].grad.numpy())
        e_b = rel_err(first[2][i].cpu().numpy(), params[i]["beta"].grad.numpy())
        print(f"full-size backward (8 x 64000 vs oracle) layer {i}: dW {e_w:.2e} dgamma {e_g:.2e} dbeta {e_b:.2e}")
        assert max(e_w, e_g, e_b) < FULL_BWD_TOL[i], (i, e_w, e_g, e_b)
