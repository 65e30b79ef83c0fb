"""Parity of the conv feature encoder kernels (layer-0 SIMT kernel, tcgen05 implicit-GEMM layers 1-6) with the
fp32 oracle.  Tolerance: 1e-2 norm-relative for bf16 features (north star); per-layer checks are tighter."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.utils import synthetic

pytestmark = pytest.mark.gpu
VARIANTS = (1, 2, 3)  # 1-CTA tiles, CTA pair splitting the channels, 2-SM UMMA pair splitting the frames (the default, 4,
                      # picks 2 or 3 per layer and runs in every test that does not select a variant)


def _layer_params(layers, dev):
    w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
    g = [None if l["gamma"] is None else torch.from_numpy(l["gamma"]).to(dev) for l in layers]
    b = [None if l["beta"] is None else torch.from_numpy(l["beta"]).to(dev) for l in layers]
    return w, g, b


@pytest.mark.parametrize("mode,l0,variant", [("layer", 0, 2), ("layer", 1, 1), ("layer", 1, 2), ("layer", 2, 2),
                                             ("layer", 3, 2), ("group", 0, 2)],
                         ids=["layer-simt", "layer-tc-v1", "layer-tc-v2", "layer-tc-fold", "layer-tc-fold-split",
                              "group-simt"])
def test_layer0_vs_oracle(dev, mode, l0, variant):
    ops.set_layer0_variant(l0)
    ops.set_frontend_variant(variant)
    layers = synthetic.frontend_weights(mode, seed=2)
    x = synthetic.waveforms(3, 4000, seed=8)[0] * 10  # roughly unit variance like a z-normed waveform
    ref = oracle.conv_frontend(torch.from_numpy(x), layers, mode, return_all=True)[0]  # [B,512,T0]
    w, g, b = _layer_params(layers, dev)
    out = ops.conv_layer0(torch.from_numpy(x).to(dev), w[0], g[0], b[0], mode)
    T, P = ops.frontend_geometry(4000)
    assert out.shape == (3, P[0], 512)
    got = out[:, :T[0]].float().transpose(1, 2).cpu().numpy()
    assert rel_err(got, ref.numpy()) < 6e-3   # bf16 output rounding (2^-9) dominates
    assert not out[:, T[0]:].any()            # pitch padding is zero-filled
    ops.set_layer0_variant(ops.DEFAULT_LAYER0_VARIANT)
    ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)


def test_layer0_tc_is_fp32_class(dev):
    """The hi/lo-split tensor-core layer 0 must agree with the SIMT fp32 kernel to bf16 output rounding (1 ulp)."""
    layers = synthetic.frontend_weights("layer", seed=6)
    x = synthetic.waveforms(4, 16000, seed=2)[0] * 10
    w, g, b = _layer_params(layers, dev)
    xd = torch.from_numpy(x).to(dev)
    ops.set_layer0_variant(0)
    ref = ops.conv_layer0(xd, w[0], g[0], b[0], "layer").float()
    ops.set_layer0_variant(1)
    got = ops.conv_layer0(xd, w[0], g[0], b[0], "layer").float()
    diff = (got - ref).abs()
    assert float(diff.max()) <= 2 ** -7 * float(ref.abs().max())      # never more than ~1 bf16 ulp of the range
    assert float((diff > 0).float().mean()) < 0.05                     # and almost always bit-identical
    for v in (2, 3):                                                   # LayerNorm folded into the GEMM operands
        ops.set_layer0_variant(v)
        fold = ops.conv_layer0(xd, w[0], g[0], b[0], "layer").float()
        diff = (fold - ref).abs()
        assert float(diff.max()) <= 2 ** -7 * float(ref.abs().max())
        assert float((diff > 0).float().mean()) < 0.05
    # DC offset 20x the signal: mean*rstd ~ 20, the folded operands must still cancel to fp32-class accuracy
    xo = xd + 2.0
    ops.set_layer0_variant(0)
    ref = ops.conv_layer0(xo, w[0], g[0], b[0], "layer").float()
    ops.set_layer0_variant(2)
    fold = ops.conv_layer0(xo, w[0], g[0], b[0], "layer").float()
    diff = (fold - ref).abs()
    assert float(diff.max()) <= 2 ** -6 * float(ref.abs().max()) and float((diff > 0).float().mean()) < 0.10
    ops.set_layer0_variant(ops.DEFAULT_LAYER0_VARIANT)


@pytest.mark.parametrize("variant", VARIANTS, ids=["v1", "v2", "v3-2sm"])
@pytest.mark.parametrize("k,rows_out,norm", [(3, 128, True), (2, 128, True), (3, 1000, True), (2, 777, False),
                                             (3, 128 * 149 + 5, True)])
def test_gemm_layer_vs_torch(dev, variant, k, rows_out, norm):
    """One implicit-GEMM layer against conv1d+LayerNorm+GELU in fp32 on the same bf16-rounded operands."""
    ops.set_frontend_variant(variant)
    rs = np.random.RandomState(k * 1000 + rows_out)
    act = torch.from_numpy(rs.standard_normal((2 * rows_out, 512)).astype(np.float32)).bfloat16()
    w = torch.from_numpy((rs.standard_normal((512, 512, k)) * np.sqrt(2.0 / (512 * k))).astype(np.float32))
    gamma = torch.from_numpy((1 + 0.1 * rs.standard_normal(512)).astype(np.float32))
    beta = torch.from_numpy((0.1 * rs.standard_normal(512)).astype(np.float32))
    wp = ops.pack_conv_weight(w.to(dev))
    assert wp.shape == (512, k * 512)
    assert torch.equal(wp.cpu().view(512, k, 512), w.bfloat16().permute(0, 2, 1).contiguous())
    out = ops.conv_layer(act.to(dev), wp, k, gamma.to(dev) if norm else None, beta.to(dev) if norm else None,
                         out_dtype=torch.float32)
    torch.cuda.synchronize()
    # reference: frames 2m .. 2m+k-1 of the (zero-extended) input, fp32 math on bf16-rounded operands
    a = torch.cat([act.float(), torch.zeros(2, 512)], 0)              # rows past the end read as zero (TMA OOB fill)
    h = F.conv1d(a.t()[None], w.bfloat16().float(), stride=2)[0].t()  # [rows_out(+), 512]
    h = h[:rows_out]
    if norm:
        h = F.layer_norm(h, (512,), gamma, beta, eps=1e-5)
    ref = F.gelu(h)
    err = rel_err(out.cpu().numpy(), ref.numpy())
    assert err < 2e-4, err
    # bf16 output path
    out16 = ops.conv_layer(act.to(dev), wp, k, gamma.to(dev) if norm else None, beta.to(dev) if norm else None)
    assert rel_err(out16.float().cpu().numpy(), ref.numpy()) < 5e-3
    ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)


@pytest.mark.parametrize("variant", VARIANTS, ids=["v1", "v2", "v3-2sm"])
@pytest.mark.parametrize("mode", ["layer", "group"])
def test_frontend_golden(dev, golden, variant, mode):
    ops.set_frontend_variant(variant)
    g = golden(f"frontend_{mode}")
    layers = synthetic.frontend_weights(mode, seed=int(g["seed"]))
    w, gm, bt = _layer_params(layers, dev)
    y = ops.conv_frontend(torch.from_numpy(g["x"]).to(dev), w, gm, bt, mode)
    assert y.shape == (2, 12, 512)
    got = y.transpose(1, 2).cpu().numpy()
    assert got.shape == g["y"].shape
    assert rel_err(got, g["y"]) < 1e-2
    ops.set_frontend_variant(ops.DEFAULT_FRONTEND_VARIANT)


@pytest.mark.parametrize("mode,B,L", [("layer", 3, 16000), ("group", 2, 16000), ("layer", 1, 400), ("layer", 2, 12345),
                                      ("layer", 2, 32000), ("layer", 1, 192000), ("group", 1, 96000), ("layer", 5, 80000)])
def test_frontend_vs_oracle(dev, mode, B, L):
    layers = synthetic.frontend_weights(mode, seed=4)
    x = synthetic.waveforms(B, L, seed=21)[0]
    x = (x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)
    ref = oracle.conv_frontend(torch.from_numpy(x), layers, mode)
    w, gm, bt = _layer_params(layers, dev)
    y = ops.conv_frontend(torch.from_numpy(x).to(dev), w, gm, bt, mode)
    assert tuple(y.shape) == (B, ref.shape[2], 512)
    err = rel_err(y.transpose(1, 2).cpu().numpy(), ref.numpy())
    assert err < 1e-2, err
    y16 = ops.conv_frontend(torch.from_numpy(x).to(dev), w, gm, bt, mode, out_dtype=torch.bfloat16)
    assert rel_err(y16.float().transpose(1, 2).cpu().numpy(), ref.numpy()) < 1.5e-2


def test_frontend_full_size_batch_independence(dev):
    """BASELINE shape 64 x 64000: every utterance's features are bit-identical to running that utterance alone
    (rows of the implicit GEMM are independent), and two sampled utterances match the fp32 oracle."""
    B, L = 64, 64000
    layers = synthetic.frontend_weights("layer", seed=0)
    x = synthetic.waveforms(B, L, seed=1234)[0]
    x = ((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype(np.float32)
    w, gm, bt = _layer_params(layers, dev)
    xd = torch.from_numpy(x).to(dev)
    y = ops.conv_frontend(xd, w, gm, bt, "layer").clone()
    assert y.shape == (B, 199, 512) and torch.isfinite(y).all()
    for b in (0, 37, 63):
        yb = ops.conv_frontend(xd[b:b + 1], w, gm, bt, "layer")
        assert torch.equal(yb[0], y[b])
    for b in (5, 63):
        ref = oracle.conv_frontend(torch.from_numpy(x[b:b + 1]), layers, "layer")
        assert rel_err(y[b].t().cpu().numpy(), ref[0].numpy()) < 1e-2


def test_forward_paths_are_deterministic(dev):
    """No atomics and fixed reduction orders in mix, conv frontend forward and loss: two runs are bit-identical."""
    clean, noise, snr_idx, table = synthetic.waveforms(6, 16000, seed=77)
    args = (torch.from_numpy(clean).to(dev), torch.from_numpy(noise).to(dev), torch.from_numpy(snr_idx).to(dev),
            [float(v) for v in table])
    c1, n1, _ = ops.mix_normalize(*args)
    c2, n2, _ = ops.mix_normalize(*args)
    assert torch.equal(c1, c2) and torch.equal(n1, n2)
    layers = synthetic.frontend_weights("layer", seed=1)
    w, g, b = _layer_params(layers, dev)
    y1 = ops.conv_frontend(n1, w, g, b, "layer").clone()
    y2 = ops.conv_frontend(n1, w, g, b, "layer")
    assert torch.equal(y1, y2)
    p, z = y1.mean(1), ops.conv_frontend(c1, w, g, b, "layer").mean(1)
    assert torch.equal(ops.byol_loss(p, z), ops.byol_loss(p, z))


def test_tile_order_is_a_pure_scheduling_knob(dev):
    """Consecutive GEMM layers walk their tiles in opposite directions by default (L2 locality); forcing every layer
    first-to-last gives bit-identical features, for the inference and the tape-writing forward."""
    layers = synthetic.frontend_weights("layer", seed=4)
    w, g, b = _layer_params(layers, dev)
    x = torch.from_numpy(synthetic.waveforms(3, 20000, seed=5)[0] * 10).to(dev)
    try:
        ops.set_tile_order(0)
        y0 = ops.conv_frontend(x, w, g, b, "layer")
        t0, _ = ops.conv_frontend_train(x, w, g, b)
        ops.set_tile_order(1)
        y1 = ops.conv_frontend(x, w, g, b, "layer")
        t1, _ = ops.conv_frontend_train(x, w, g, b)
    finally:
        ops.set_tile_order(1)
    assert torch.equal(y0, y1) and torch.equal(t0, t1)


def test_launch_overlap_and_l2_policy_do_not_change_results(dev):
    """Programmatic dependent launch (set-up of a layer's kernel overlaps the previous kernel's tail) and the evict-first
    policy of the output stores are performance features: switching them off through the timing hooks (NRSE_EXPERIMENT 512 /
    8, read once per process) must give bit-identical features.  A race between consecutive layers would show up here."""
    import hashlib, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, hashlib, torch; sys.path.insert(0, %r);"
        "from nrse_b200 import ops; from nrse_b200.utils import synthetic;"
        "d = torch.device('cuda:0'); L = synthetic.frontend_weights('layer', seed=4);"
        "x = torch.from_numpy(synthetic.waveforms(24, 48000, seed=3)[0]).to(d) * 10;"
        "w = [torch.from_numpy(l['conv']).to(d) for l in L]; g = [torch.from_numpy(l['gamma']).to(d) for l in L];"
        "b = [torch.from_numpy(l['beta']).to(d) for l in L];"
        "h = hashlib.sha256();"
        "[h.update(ops.conv_frontend(x, w, g, b, 'layer', out_dtype=torch.bfloat16).view(torch.int16).cpu().numpy().tobytes())"
        " for _ in range(3)];"
        "print('SHA', h.hexdigest())" % root)
    digests = {}
    for flags in (0, 8, 512, 520):
        env = dict(os.environ, NRSE_EXPERIMENT=str(flags))
        if flags == 0:
            env.pop("NRSE_EXPERIMENT")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[flags] = [l for l in r.stdout.splitlines() if l.startswith("SHA")][-1]
    assert len(set(digests.values())) == 1, digests
