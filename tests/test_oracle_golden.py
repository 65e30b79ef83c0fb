"""Pin the oracle against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import torch

import oracle
from nrse_b200.utils import synthetic
from conftest import rel_err

torch.set_num_threads(1)


def test_mix_byol_matches_reference_dataset(golden):
    g = golden("mix_byol")
    c, n, st = oracle.mix_normalize_batch(g["clean"], g["noise"], g["snr_idx"], g["snr_table"], peak_norm=True)
    assert st.tolist() == [0] * len(st)
    # same ops in the same order on the same thread count -> bit-identical
    assert np.array_equal(c.numpy(), g["clean_out"])
    assert np.array_equal(n.numpy(), g["noisy_out"])


def test_mix_emotion_matches_reference(golden):
    g = golden("mix_emotion")
    for rows, noise in ((slice(0, 2), g["noise_short"]), (slice(2, 4), g["noise_long"])):
        _, n, st = oracle.mix_normalize_batch(g["clean"][rows], noise, g["snr_idx"][rows], g["snr_table"],
                                              peak_norm=False)
        assert st.tolist() == [0, 0]
        assert np.array_equal(n.numpy(), g["noisy_out"][rows])


def test_mix_edge_none_conditions(golden):
    g = golden("mix_edge")
    expect = [oracle.mix.STATUS_SPEECH_NAN, oracle.mix.STATUS_NOISE_NAN, oracle.mix.STATUS_SPEECH_POWER,
              oracle.mix.STATUS_NOISE_POWER, oracle.mix.STATUS_SCALED_NOISE_NAN, oracle.mix.STATUS_SCALE_INVALID, 0, 0]
    for b in range(8):
        r = oracle.add_noise_to_speech(torch.from_numpy(g["clean"][b:b + 1]), torch.from_numpy(g["noise"][b:b + 1]), 10)
        assert (r is None) == bool(g["is_none"][b])
        _, _, st = oracle.mix_normalize_item(torch.from_numpy(g["clean"][b:b + 1]),
                                             torch.from_numpy(g["noise"][b:b + 1]), 10)
        assert st == expect[b]


def test_byol_loss_matches_reference(golden):
    g = golden("byol_loss")
    for name in ("b64", "b2", "b5_d96"):
        p = torch.from_numpy(g[f"{name}_p"]).requires_grad_(True)
        loss = oracle.byol_loss(p, torch.from_numpy(g[f"{name}_z"]))
        loss.backward()
        assert np.array_equal(loss.detach().numpy(), g[f"{name}_loss"])
        assert np.array_equal(p.grad.numpy(), g[f"{name}_grad"])


def test_ema_matches_reference(golden):
    g = golden("ema")
    n = int(g["n"])
    online = [torch.from_numpy(g[f"o{i}"]) for i in range(n)]
    for decay in (0.996, 0.997):
        target = [torch.from_numpy(g[f"t{i}"]) for i in range(n)]
        target = oracle.ema_update(online, target, decay)
        target = oracle.ema_update(online, target, decay)
        for i in range(n):
            assert np.array_equal(target[i].numpy(), g[f"r{int(decay * 1000)}_{i}"])


def test_frontend_matches_reference(golden):
    for mode in ("layer", "group"):
        g = golden(f"frontend_{mode}")
        layers = synthetic.frontend_weights(mode, seed=int(g["seed"]))
        outs = oracle.conv_frontend(torch.from_numpy(g["x"]), layers, mode, return_all=True)
        assert list(outs[-1].shape) == list(g["y"].shape)
        assert rel_err(outs[-1].numpy(), g["y"]) < 1e-6
        assert rel_err(outs[0].numpy()[:, :, :64], g["y0"]) < 1e-6
        assert rel_err(outs[1].numpy()[:, :, :64], g["y1"]) < 1e-6


def test_frontend_matches_installed_transformers():
    """The restatement vs the third-party classes themselves (present on the GPU box as well)."""
    from transformers import WavLMConfig
    from transformers.models.wavlm.modeling_wavlm import WavLMFeatureEncoder

    for mode in ("layer", "group"):
        layers = synthetic.frontend_weights(mode, seed=3)
        fe = WavLMFeatureEncoder(WavLMConfig(feat_extract_norm=mode, conv_bias=False)).eval()
        with torch.no_grad():
            for i, cl in enumerate(fe.conv_layers):
                cl.conv.weight.copy_(torch.from_numpy(layers[i]["conv"]))
                if layers[i]["gamma"] is not None:
                    cl.layer_norm.weight.copy_(torch.from_numpy(layers[i]["gamma"]))
                    cl.layer_norm.bias.copy_(torch.from_numpy(layers[i]["beta"]))
            x = torch.from_numpy(synthetic.waveforms(2, 2000, seed=5)[0]) * 10
            ref = fe(x)
        got = oracle.conv_frontend(x, layers, mode)
        assert rel_err(got.numpy(), ref.numpy()) < 1e-6
        assert got.shape[-1] == synthetic.conv_out_lengths(2000)[-1] == oracle.conv_out_lengths(2000)[-1]


def test_synthetic_is_deterministic():
    a = synthetic.waveforms(3, 1000, seed=9)
    b = synthetic.waveforms(3, 1000, seed=9)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert synthetic.conv_out_lengths(64000) == [12799, 6399, 3199, 1599, 799, 399, 199]


def test_single_mufu_gelu_fit_is_exact_gelu():
    """The CUDA epilogue evaluates erf-GELU in the halved argument w = x/2, a = |w|: gelu = w + a*(1 - 2^-q(a)) with a
    polynomial fit q of -log2 erfc(sqrt2 a) (csrc/frontend.cu::gelu2h; cubic by default, quintic with
    -DNRSE_GELU_DEG=5).  Pin that formula, evaluated in float32 on the CPU, against torch's exact GELU: the error must
    stay far below one bf16 ulp (4e-3 at |x| ~ 1).  The derivative fit of the backward kernel is pinned below."""
    x = np.linspace(-12, 12, 400001).astype(np.float32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    ref = torch.nn.functional.gelu(xt)
    ref.sum().backward()
    fits = {3: ([2.281832695007324, 1.9530425071716309, 0.2210327833890915], 1e-4),
            5: ([2.3020009994506836, 1.838383436203003, 0.4171730577945709, -0.11517950147390366,
                 0.015619270503520966], 2e-6)}
    for deg, (coef, tol) in fits.items():
        c = np.array(coef, dtype=np.float32)
        w = x * np.float32(0.5)
        a = np.abs(w)
        p = np.full_like(a, c[-1])
        for k in range(deg - 2, -1, -1):
            p = p * a + c[k]
        q = p * a
        assert np.all(np.diff(q[x >= 0]) >= 0) and q[-1] > 126          # 2^-q flushes to 0 by itself: no clamp
        e = np.exp2(-q.astype(np.float64)).astype(np.float32)             # erfc(sqrt2 a)
        gelu = w + a * (np.float32(1) - e)
        assert np.abs(gelu - ref.detach().numpy()).max() < tol, deg
    u = np.minimum(np.abs(x), np.float32(5.6))
    # derivative (csrc/frontend.cu::gelu_grad2): gelu'(v) = 0.5 + copysign(0.5 - exp(-v^2/2) r(|v|), v), degree-6 fit r
    k = np.float32(0.3989422804)
    rc = [np.float32(k * v) for v in (1.2532488107681274, -1.9976462125778198, 0.6104238033294678, -0.2893761098384857,
                                      0.09663444012403488, -0.019207235425710678, 0.0016475850716233253)]
    r = np.zeros_like(u)
    for coef in reversed(rc):
        r = r * u + coef
    ex = np.exp2(u * u * np.float32(-0.72134752)).astype(np.float32)
    grad = np.float32(0.5) + np.copysign(np.float32(0.5) - ex * r, x)
    assert np.abs(grad - xt.grad.numpy()).max() < 5e-5                    # bf16 resolution is 4e-3


def _optim_fixture(g):
    names = [str(n) for n in g["names"]]
    no_grad = {str(n) for n in g["no_grad"]}
    n = len(names)
    params = [torch.from_numpy(g[f"p0_{i}"].copy()) for i in range(n)]
    targets = [torch.from_numpy(g[f"t0_{i}"].copy()) if bool(g[f"has_twin_{i}"]) else None for i in range(n)]
    has_grad = [names[i] not in no_grad for i in range(n)]

    def grads(k):
        return [torch.from_numpy(synthetic.optim_step_grad(int(g["seed"]), k, i, tuple(params[i].shape),
                                                           float(g["grad_scales"][k]))) if has_grad[i] else None
                for i in range(n)]
    return names, params, targets, has_grad, grads


def test_optim_step_matches_torch_adamw_and_reference_ema(golden):
    """oracle.clip_adamw_ema_step == clip_grad_norm_ + torch.optim.AdamW + reference EMA, three steps, bit for bit."""
    g = golden("optim_step")
    names, params, targets, has_grad, grads = _optim_fixture(g)
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for k in range(int(g["steps"])):
        norm = oracle.clip_adamw_ema_step(params, grads(k), m, v, targets, k + 1, float(g["lr"]),
                                          weight_decay=float(g["weight_decay"]), max_norm=1.0,
                                          ema_decay=float(g["ema_decay"]))
        assert abs(float(norm) - float(g["norms"][k])) <= 1e-6 * float(g["norms"][k])
    for i in range(len(names)):
        assert np.array_equal(params[i].numpy(), g[f"p_{i}"]), names[i]
        if targets[i] is not None:
            assert np.array_equal(targets[i].numpy(), g[f"t_{i}"]), names[i]
        if has_grad[i]:
            assert np.array_equal(m[i].numpy(), g[f"m_{i}"]) and np.array_equal(v[i].numpy(), g[f"v_{i}"]), names[i]


def _emotion_mask(g):
    lens = [int(v) for v in g["mask_lens"]]
    mask = torch.zeros(len(lens), max(lens))
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    return mask


def test_attentive_pooling_and_ccc_match_reference(golden):
    g = golden("emotion")
    mask = _emotion_mask(g)
    assert oracle.compute_length_from_mask(mask) == [int(v) for v in g["feat_lens"]]
    for xs_key, out_key, dx_key in (("xs", "out", "dxs"), (None, "out_clamped", "dxs_clamped")):
        xs = torch.from_numpy(g["xs"].copy())
        if xs_key is None:
            xs[:, :, 3] = 0.25
            xs[:, :, 7] = 1e-3 * xs[:, :, 7]
        xs.requires_grad_(True)
        w = torch.from_numpy(g["sap_w"]).requires_grad_(True)
        b = torch.from_numpy(g["sap_b"]).requires_grad_(True)
        a = torch.from_numpy(g["attention"]).requires_grad_(True)
        out = oracle.attentive_statistics_pooling(xs, mask, w, b, a)
        out.backward(torch.from_numpy(g["gout"]))
        assert np.array_equal(out.detach().numpy(), g[out_key])
        assert np.array_equal(xs.grad.numpy(), g[dx_key])
        if xs_key is not None:
            assert np.array_equal(w.grad.numpy(), g["d_sap_w"]) and np.array_equal(a.grad.numpy(), g["d_attention"])
    pred = torch.from_numpy(g["ccc_pred"]).requires_grad_(True)
    loss = oracle.ccc_loss(pred, torch.from_numpy(g["ccc_targ"]))
    loss.backward()
    assert np.array_equal(loss.detach().numpy(), g["ccc_loss"]) and np.array_equal(pred.grad.numpy(), g["ccc_grad"])
    # the product's vectorised ccc_loss is plain torch: check it here on the CPU as well
    from nrse_b200.train import ccc_loss, compute_ccc
    pred2 = torch.from_numpy(g["ccc_pred"]).requires_grad_(True)
    loss2 = ccc_loss(pred2, torch.from_numpy(g["ccc_targ"]))
    loss2.backward()
    assert rel_err(loss2.detach().numpy(), g["ccc_loss"]) < 1e-6 and rel_err(pred2.grad.numpy(), g["ccc_grad"]) < 1e-5
    assert float(ccc_loss(pred2[:1], torch.from_numpy(g["ccc_targ"][:1]))) == 0.0
    c = compute_ccc(g["ccc_pred"][:, 0], g["ccc_targ"][:, 0])
    assert abs((1 - c) - float(oracle.ccc_loss(torch.from_numpy(g["ccc_pred"][:, :1]),
                                               torch.from_numpy(g["ccc_targ"][:, :1])))) < 1e-6


def featproj_wavlm_config():
    """The configuration tests/golden/make_golden.py::gen_feature_projection used."""
    from transformers import WavLMConfig
    cfg = WavLMConfig(hidden_size=1024, num_hidden_layers=1, num_attention_heads=16, intermediate_size=64,
                      feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False,
                      num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout", "feat_proj_dropout", "final_dropout",
              "layerdrop", "mask_time_prob", "mask_feature_prob"):
        setattr(cfg, k, 0.0)
    cfg.apply_spec_augment = False
    return cfg


def test_feature_projection_fixture_matches_installed_transformers(golden):
    """The fixture written through the reference's WavLMEncoder equals the installed transformers classes on the same
    seeded weights (the GPU box has the same transformers: the seeded init is reproducible there)."""
    from transformers import WavLMModel
    g = golden("feature_projection")
    torch.manual_seed(int(g["seed"]))
    model = WavLMModel(featproj_wavlm_config()).train()
    out = model(torch.from_numpy(g["x"]))
    assert rel_err(out.last_hidden_state.detach().numpy(), g["last_hidden"]) < 1e-5
    assert rel_err(out.extract_features.detach().numpy(), g["norm"]) < 1e-5


def test_mix_batch_attempts_restates_the_attempt_loop():
    """``oracle.mix_batch_attempts`` (ref:src/data/noisy_speech_dataset.py:55-149 with explicit donors): a healthy batch is the
    plain per-row chain; a row whose noise crop is unusable takes the next usable donor's noise AND SNR draw; a row whose
    clean crop is silent never recovers and is substituted by the nearest following good row (or stays zero-filled)."""
    import numpy as np
    import torch

    import oracle
    from nrse_b200.utils import synthetic
    B, L = 6, 2000
    clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=7)
    c0, n0, st0 = oracle.mix_normalize_batch(clean, noise, snr_idx, table)
    c, n, st, used, rejected = oracle.mix_batch_attempts(clean, noise, snr_idx, table)
    assert torch.equal(c, c0) and torch.equal(n, n0) and st.tolist() == st0.tolist() == [0] * B
    assert used.tolist() == snr_idx.tolist() and rejected == 0
    noise[1] = 0.0          # rows 1 and 2 have unusable noise: row 1 needs two retries (donors 2, 3), row 2 one (donor 3)
    noise[2] = np.nan
    clean[5] = 0.0          # silent clean crop: never recovers; last row wraps around to row 0
    c, n, st, used, rejected = oracle.mix_batch_attempts(clean, noise, snr_idx, table)
    assert st.tolist() == [0, 0, 0, 0, 0, 3] and rejected == 1
    for b in (1, 2):
        cr, nr, sr = oracle.mix_normalize_batch(clean[b:b + 1], noise[3:4], snr_idx[3:4], table)
        assert sr.tolist() == [0] and torch.equal(n[b], nr[0]) and torch.equal(c[b], cr[0]) and used[b] == snr_idx[3]
    assert torch.equal(c[5], c[0]) and torch.equal(n[5], n[0]) and used[5] == snr_idx[0]
    c2, n2, st2, _, rejected2 = oracle.mix_batch_attempts(clean, noise, snr_idx, table, max_attempts=2, substitute=False)
    assert st2.tolist() == [0, 2, 0, 0, 0, 3] and rejected2 == 2 and not c2[1].any() and not n2[5].any()
