#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

The reference's own modules are imported from /root/reference via sys.path; nothing is
copied.  Shims, all explicit below:
  1. ``AutoModel.from_pretrained`` -> random-init ``WavLMModel(WavLMConfig(...))`` (no weights
     offline); the conv feature-encoder parameters are then overwritten with
     ``nrse_b200.utils.synthetic.frontend_weights`` so the GPU box can rebuild them from a seed.
  2. ``AutoFeatureExtractor`` -> ``Wav2Vec2FeatureExtractor(do_normalize=True)`` (wavlm-large's
     preprocessor settings; only ``do_normalize`` affects the path).
  3. ``load_and_process_audio`` -> in-memory tensors (torchaudio cannot decode in this image).
Outputs (float arrays kept small; inputs are stored too so fixtures are self-contained):
  mix_byol.npz, mix_emotion.npz, mix_edge.npz, byol_loss.npz, ema.npz, frontend_layer.npz,
  frontend_group.npz, byol_step.npz, byol_state_dict_keys.json, optim_step.npz,
  emotion.npz, emotion_state_dict_keys.json, evaluate_byol.npz, feature_projection.npz
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from nrse_b200.utils import synthetic  # noqa: E402

import transformers  # noqa: E402
from transformers import Wav2Vec2FeatureExtractor, WavLMConfig, WavLMModel  # noqa: E402

# --- reference modules, unmodified -------------------------------------------------------------
import src.models.encoder as ref_encoder_mod  # noqa: E402
import src.data.noisy_speech_dataset as ref_ds_mod  # noqa: E402
from src.data.augment import add_noise_to_speech as ref_add_noise  # noqa: E402
from src.models.byol import BYOLSpeechModel as RefBYOL, byol_loss as ref_byol_loss  # noqa: E402

torch.set_num_threads(1)  # single-thread reductions: the most reproducible summation order


def small_wavlm_config(norm_mode):
    """Full-size conv feature encoder (that is the path); tiny transformer (out of scope)."""
    return WavLMConfig(
        hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
        feat_extract_norm=norm_mode, do_stable_layer_norm=(norm_mode == "layer"), conv_bias=False,
        num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4,
    )


def feature_extractor():
    return Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0,
                                    do_normalize=True, return_attention_mask=True)


# ------------------------------------------------------------------------------------------------
def gen_mix_byol(B=6, L=8000, Ln=8000, seed=11, snr_range=(2, 5, 10, 15, 20)):
    """The real ``NoiseRobustSpeechDataset.__getitem__`` with an in-memory loader."""
    clean, noise, _, _ = synthetic.waveforms(B, L, seed=seed, n_noise=Ln, snr_range=snr_range)
    tmp = tempfile.mkdtemp()
    cdir, ndir = os.path.join(tmp, "clean"), os.path.join(tmp, "noise")
    os.makedirs(cdir), os.makedirs(ndir)
    for b in range(B):
        open(os.path.join(cdir, f"c{b:03d}.wav"), "w").close()
        open(os.path.join(ndir, f"n{b:03d}.wav"), "w").close()
    ds = ref_ds_mod.NoiseRobustSpeechDataset(cdir, ndir, sample_rate=16000, max_audio_length=L / 16000,
                                             snr_range=list(snr_range), feature_extractor=feature_extractor())
    ds.clean_files.sort(), ds.noise_files.sort()
    picked = {}

    def loader(path):
        name = os.path.basename(path)
        idx = int(name[1:4])
        if name[0] == "n":
            picked["noise"] = idx
            return torch.from_numpy(noise[idx:idx + 1].copy())
        return torch.from_numpy(clean[idx:idx + 1].copy())

    ds._load_and_process_audio = loader
    outs_c, outs_n, snrs, nidx = [], [], [], []
    random.seed(seed)
    for b in range(B):
        item = ds[b]
        outs_c.append(item["clean_input_values"].numpy()[0])
        outs_n.append(item["noisy_input_values"].numpy()[0])
        snrs.append(item["snr"])
        nidx.append(picked["noise"])
    snr_idx = np.array([snr_range.index(s) for s in snrs], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "mix_byol.npz"),
                        clean=clean, noise=noise[np.array(nidx)], snr_idx=snr_idx,
                        snr_table=np.array(snr_range, dtype=np.float64),
                        clean_out=np.stack(outs_c), noisy_out=np.stack(outs_n))
    print("mix_byol", np.stack(outs_c).shape, snrs, nidx)


def gen_mix_emotion(B=4, L=6000, seed=12, snr_range=(4, 8)):
    """Emotion fine-tune variant (ref:src/data/emotion_dataset.py:177-203): the reference's
    ``add_noise_to_speech`` then the HF extractor, no peak-norm; noise shorter than speech (tiled)
    on rows 0-1 and longer (truncated) on rows 2-3."""
    fe = feature_extractor()
    clean, _, snr_idx, _ = synthetic.waveforms(B, L, seed=seed, snr_range=snr_range)
    noise_short = synthetic.waveforms(B, 2500, seed=seed + 1)[1]
    noise_long = synthetic.waveforms(B, 7000, seed=seed + 2)[1]
    outs, noises = [], []
    for b in range(B):
        nz = noise_short[b:b + 1] if b < 2 else noise_long[b:b + 1]
        wav = torch.from_numpy(clean[b:b + 1].copy())
        noisy = ref_add_noise(wav, torch.from_numpy(nz.copy()), snr_range[int(snr_idx[b])])
        assert noisy is not None
        inputs = fe(noisy.squeeze().numpy(), sampling_rate=16000, return_tensors="pt")
        outs.append(inputs.input_values.squeeze(0).numpy())
        noises.append(nz[0])
    np.savez_compressed(os.path.join(HERE, "mix_emotion.npz"), clean=clean,
                        noise_short=np.stack(noises[:2]), noise_long=np.stack(noises[2:]),
                        snr_idx=snr_idx, snr_table=np.array(snr_range, dtype=np.float64),
                        noisy_out=np.stack(outs))
    print("mix_emotion", np.stack(outs).shape)


def gen_mix_edge(L=4000, seed=13):
    """``None`` exits of the reference's ``add_noise_to_speech`` (ref:src/data/augment.py:7-64)."""
    clean, noise, _, _ = synthetic.waveforms(8, L, seed=seed)
    clean[0, 17] = np.nan                      # speech NaN         -> None
    noise[1, 5] = np.nan                       # noise NaN          -> None
    clean[2] *= 1e-6                           # speech power<1e-10 -> None
    noise[3] *= 1e-6                           # noise power<1e-10  -> None
    noise[4, 9] = np.inf                       # Pn=inf -> scale=0 -> inf*0=NaN -> None
    clean[5, 3] = np.inf                       # Ps=inf -> scale inf -> None
    # rows 6,7 are valid
    is_none = []
    for b in range(8):
        r = ref_add_noise(torch.from_numpy(clean[b:b + 1].copy()), torch.from_numpy(noise[b:b + 1].copy()), 10)
        is_none.append(r is None)
    np.savez_compressed(os.path.join(HERE, "mix_edge.npz"), clean=clean, noise=noise,
                        is_none=np.array(is_none))
    print("mix_edge none:", is_none)


def gen_byol_loss(seed=21):
    cases = {}
    for name, (B, D) in {"b64": (64, 1024), "b2": (2, 1024), "b5_d96": (5, 96)}.items():
        p, z = synthetic.embeddings(B, D, seed=seed + B)
        if name == "b64":
            z[0] = p[0]            # identical -> sim 1
            z[1] = -p[1]           # opposite  -> sim -1
            p[2] = 0.0             # all-zero online row (the +1e-10 path)
            z[3] = 0.0             # all-zero target row
        pt = torch.from_numpy(p.copy()).requires_grad_(True)
        loss = ref_byol_loss(pt, torch.from_numpy(z.copy()))
        loss.backward()
        cases[f"{name}_p"], cases[f"{name}_z"] = p, z
        cases[f"{name}_loss"] = loss.detach().numpy()
        cases[f"{name}_grad"] = pt.grad.numpy()
        print("byol_loss", name, float(loss))
    np.savez_compressed(os.path.join(HERE, "byol_loss.npz"), **cases)


def gen_ema(seed=31):
    """``BYOLSpeechModel._update_target_network`` of the reference on a shimmed (tiny) WavLM."""
    cfg_small = WavLMConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64,
                            conv_dim=(16,) * 7, num_conv_pos_embeddings=8, num_conv_pos_embedding_groups=2,
                            feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    orig = ref_encoder_mod.AutoModel.from_pretrained
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg_small))
    try:
        torch.manual_seed(seed)
        model = RefBYOL({"model": {"name": "shim", "projection_dim": 24, "prediction_dim": 48, "ema_decay": 0.996}})
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig
    rs = np.random.RandomState(seed)
    out = {}
    with torch.no_grad():
        for mod in (model.online_encoder, model.online_projector):
            for p in mod.parameters():
                p.add_(torch.from_numpy((0.05 * rs.standard_normal(tuple(p.shape))).astype(np.float32)))
    pairs = list(zip(model.online_encoder.parameters(), model.target_encoder.parameters())) + \
        list(zip(model.online_projector.parameters(), model.target_projector.parameters()))
    for i, (o, t) in enumerate(pairs):
        out[f"o{i}"] = o.detach().numpy().copy()
        out[f"t{i}"] = t.detach().numpy().copy()
    for decay in (0.996, 0.997):
        model.ema_decay = decay
        for i, (o, t) in enumerate(pairs):       # reset targets
            t.data = torch.from_numpy(out[f"t{i}"].copy())
        model._update_target_network()
        model._update_target_network()            # two consecutive steps
        for i, (o, t) in enumerate(pairs):
            out[f"r{int(decay * 1000)}_{i}"] = t.detach().numpy().copy()
    out["n"] = np.array(len(pairs))
    np.savez_compressed(os.path.join(HERE, "ema.npz"), **out)
    print("ema tensors", len(pairs), "elems", sum(o.numel() for o, _ in pairs))


def gen_frontend(norm_mode, B=2, L=4000, seed=41):
    """conv feature encoder through the reference's ``WavLMEncoder`` wrapper object
    (ref:src/models/encoder.py:6-15): ``encoder.model.feature_extractor``."""
    cfg = small_wavlm_config(norm_mode)
    orig = ref_encoder_mod.AutoModel.from_pretrained
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg))
    try:
        torch.manual_seed(seed)
        enc = ref_encoder_mod.WavLMEncoder("shim")
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig
    enc.eval()
    layers = synthetic.frontend_weights(norm_mode, seed=seed)
    with torch.no_grad():
        for i, layer in enumerate(enc.model.feature_extractor.conv_layers):
            layer.conv.weight.copy_(torch.from_numpy(layers[i]["conv"]))
            if layers[i]["gamma"] is not None:
                layer.layer_norm.weight.copy_(torch.from_numpy(layers[i]["gamma"]))
                layer.layer_norm.bias.copy_(torch.from_numpy(layers[i]["beta"]))
    # z-normalised waveform input, as the path feeds it
    clean, noise, _, _ = synthetic.waveforms(B, L, seed=seed + 1)
    x = np.stack([(c - c.mean()) / np.sqrt(c.var() + 1e-7) for c in clean]).astype(np.float32)
    with torch.no_grad():
        h = torch.from_numpy(x)[:, None]
        per_layer = []
        for conv_layer in enc.model.feature_extractor.conv_layers:
            h = conv_layer(h)
            per_layer.append(h.numpy().copy())
        y = enc.model.feature_extractor(torch.from_numpy(x)).numpy()
    assert np.array_equal(y, per_layer[-1])
    np.savez_compressed(os.path.join(HERE, f"frontend_{norm_mode}.npz"), x=x, seed=np.array(seed),
                        y=y, y0=per_layer[0][:, :, :64], y1=per_layer[1][:, :, :64])
    print("frontend", norm_mode, y.shape, float(np.abs(y).mean()))


def gen_byol_step(seed=51, B=4, L=4000):
    """The reference's BYOLSpeechModel (+ the mean-pool shim of SURVEY.md fact 3, applied to the reference's own
    ``WavLMEncoder.forward``) on a tiny transformer with the FULL-SIZE conv feature encoder: state-dict keys/shapes,
    an eval-mode forward + byol_loss, and one deterministic training step (train mode with dropout / layerdrop /
    SpecAugment disabled in the config): loss, clipped-gradient AdamW update, EMA (ref:train_byol.py:56-74)."""
    import json
    cfg = small_wavlm_config("layer")
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout", "feat_proj_dropout", "final_dropout",
              "layerdrop", "mask_time_prob", "mask_feature_prob"):
        setattr(cfg, k, 0.0)
    cfg.apply_spec_augment = False
    orig_from = ref_encoder_mod.AutoModel.from_pretrained
    orig_fwd = ref_encoder_mod.WavLMEncoder.forward
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg))
    ref_encoder_mod.WavLMEncoder.forward = lambda self, x, attention_mask=None: orig_fwd(self, x, attention_mask).mean(dim=1)
    try:
        torch.manual_seed(seed)
        model = RefBYOL({"model": {"name": "shim", "projection_dim": 96, "prediction_dim": 128, "ema_decay": 0.99}})
        keys = {k: list(v.shape) for k, v in model.state_dict().items()}
        with open(os.path.join(HERE, "byol_state_dict_keys.json"), "w") as f:
            json.dump(keys, f, indent=0, sort_keys=True)
        clean, noise, snr_idx, table = synthetic.waveforms(B, L, seed=seed)
        import oracle
        c, n, st = oracle.mix_normalize_batch(clean, noise, snr_idx, table, peak_norm=True)
        assert not st.any()
        c, n = c[:, None], n[:, None]
        out = {"clean_in": c.numpy(), "noisy_in": n.numpy(), "seed": np.array(seed)}
        model.eval()
        with torch.no_grad():
            op, tp = model(c, n)
            out["eval_online_pred"], out["eval_target_proj"] = op.numpy(), tp.numpy()
            out["eval_loss"] = ref_byol_loss(op, tp).numpy()
            out["eval_online_emb"] = model.online_encoder(c).numpy()
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
        op, tp = model(c, n)
        loss = ref_byol_loss(op, tp)
        opt.zero_grad()
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        model._update_target_network()
        out["train_loss"] = loss.detach().numpy()
        out["train_grad_norm"] = gnorm.numpy()
        sd = model.state_dict()
        for k in ("online_encoder.model.feature_extractor.conv_layers.0.conv.weight",
                  "online_encoder.model.feature_extractor.conv_layers.3.conv.weight",
                  "online_encoder.model.feature_extractor.conv_layers.6.layer_norm.weight",
                  "target_encoder.model.feature_extractor.conv_layers.3.conv.weight",
                  "target_encoder.model.encoder.layers.1.feed_forward.output_dense.weight",
                  "target_projector.layers.3.weight", "online_predictor.layers.6.bias"):
            out["after::" + k] = sd[k].numpy().reshape(-1)[:2048].copy()
        np.savez_compressed(os.path.join(HERE, "byol_step.npz"), **out)
        print("byol_step: keys", len(keys), "eval loss", float(out["eval_loss"]), "train loss", float(loss),
              "grad norm", float(gnorm))
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig_from
        ref_encoder_mod.WavLMEncoder.forward = orig_fwd


def gen_evaluate_byol(seed=81, n_batches=3, B=4, L=4000):
    """The reference's OWN ``evaluate_embedding_similarity`` and ``validate_model`` (ref:evaluate_byol.py:12-123, imported
    unmodified; matplotlib / tqdm are only imported at its module level: stubbed when absent) on the shimmed BYOL model of
    ``gen_byol_step`` and an in-memory "loader" of reference-format batches: per-SNR mean cosine similarity of the clean /
    noisy embeddings, validation loss, metrics dict."""
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import evaluate_byol as ref_eval
    cfg = small_wavlm_config("layer")
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout", "feat_proj_dropout", "final_dropout",
              "layerdrop", "mask_time_prob", "mask_feature_prob"):
        setattr(cfg, k, 0.0)
    cfg.apply_spec_augment = False
    orig_from = ref_encoder_mod.AutoModel.from_pretrained
    orig_fwd = ref_encoder_mod.WavLMEncoder.forward
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg))
    ref_encoder_mod.WavLMEncoder.forward = lambda self, x, attention_mask=None: orig_fwd(self, x, attention_mask).mean(dim=1)
    try:
        torch.manual_seed(seed)
        model = RefBYOL({"model": {"name": "shim", "projection_dim": 96, "prediction_dim": 128, "ema_decay": 0.99}})
        # a trained model has diverged online / target branches: perturb the online branch so that the similarity and the
        # loss are not the degenerate values of two identical networks
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed)
            for p in model.online_encoder.parameters():
                p.add_(0.02 * p.abs().mean() * torch.randn(p.shape, generator=g))
        import oracle
        snr_range = [2, 5, 10, 15]   # 15 never drawn below: the reference reports 0 for an empty bucket
        out = {"seed": np.array(seed), "snr_range": np.array(snr_range)}
        loader = []
        for i in range(n_batches):
            clean, noise, snr_idx, _ = synthetic.waveforms(B, L, seed=seed + 1 + i, snr_range=snr_range[:3])
            c, n, st = oracle.mix_normalize_batch(clean, noise, snr_idx, np.asarray(snr_range[:3], dtype=np.float64),
                                                  peak_norm=True)
            assert not st.any()
            snr = torch.tensor([snr_range[j] for j in snr_idx], dtype=torch.int64)
            loader.append({"clean_input_values": c[:, None], "noisy_input_values": n[:, None], "snr": snr})
            out[f"clean_{i}"], out[f"noisy_{i}"], out[f"snr_{i}"] = c.numpy(), n.numpy(), snr.numpy()
        config = {"data": {"snr_range": snr_range}}
        sims = ref_eval.evaluate_embedding_similarity(model, loader, torch.device("cpu"), config)
        val_loss, metrics = ref_eval.validate_model(model, loader, torch.device("cpu"), config)
        assert metrics["val_similarities"] == sims
        out["similarities"] = np.array([sims[s] for s in snr_range], dtype=np.float64)
        out["val_loss"] = np.array(val_loss, dtype=np.float64)
        out["val_avg_similarity"] = np.array(metrics["val_avg_similarity"], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "evaluate_byol.npz"), **out)
        print("evaluate_byol: similarities", sims, "val_loss", val_loss)
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig_from
        ref_encoder_mod.WavLMEncoder.forward = orig_fwd


def featproj_wavlm_config():
    """wavlm-large's conv stack AND feature projection (512 -> 1024); one tiny transformer layer behind it."""
    cfg = WavLMConfig(hidden_size=1024, num_hidden_layers=1, num_attention_heads=16, intermediate_size=64,
                      feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False,
                      num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout", "feat_proj_dropout", "final_dropout",
              "layerdrop", "mask_time_prob", "mask_feature_prob"):
        setattr(cfg, k, 0.0)
    cfg.apply_spec_augment = False
    return cfg


def gen_feature_projection(seed=91, B=2, L=6000):
    """Through the reference's ``WavLMEncoder`` (ref:src/models/encoder.py:5-32): the outputs of
    ``model.feature_projection`` (hf:models/wavlm/modeling_wavlm.py:93-105; captured with a forward hook), the encoder's
    ``last_hidden_state``, and the autograd gradients of the feature projection's parameters and of the last conv
    layer's weight for the loss sum(last_hidden_state * G)."""
    cfg = featproj_wavlm_config()
    orig = ref_encoder_mod.AutoModel.from_pretrained
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg))
    try:
        torch.manual_seed(seed)
        enc = ref_encoder_mod.WavLMEncoder("shim")
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig
    x = synthetic.waveforms(B, L, seed=seed)[0]
    x = ((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype(np.float32)
    captured = {}
    h = enc.model.feature_projection.register_forward_hook(lambda m, i, o: captured.update(hidden=o[0], norm=o[1]))
    enc.train()  # dropout / layerdrop / SpecAugment are zero in this config: deterministic
    for p in enc.parameters():
        p.requires_grad_(True)
    y = enc(torch.from_numpy(x)[:, None])
    h.remove()
    G = torch.from_numpy(np.random.RandomState(seed).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * G).sum().backward()
    fp = enc.model.feature_projection
    out = {"x": x, "seed": np.array(seed), "G": G.numpy(), "hidden": captured["hidden"].detach().numpy(),
           "norm": captured["norm"].detach().numpy(), "last_hidden": y.detach().numpy(),
           "d_ln_weight": fp.layer_norm.weight.grad.numpy(), "d_ln_bias": fp.layer_norm.bias.grad.numpy(),
           "d_proj_bias": fp.projection.bias.grad.numpy(),
           "d_proj_weight_rows": fp.projection.weight.grad.numpy()[::64].copy(),      # 16 of the 1024 rows
           "d_conv6_weight_rows": enc.model.feature_extractor.conv_layers[6].conv.weight.grad.numpy()[::64].copy(),
           "d_conv0_weight": enc.model.feature_extractor.conv_layers[0].conv.weight.grad.numpy()}
    np.savez_compressed(os.path.join(HERE, "feature_projection.npz"), **out)
    print("feature_projection: hidden", tuple(captured["hidden"].shape), "last_hidden", tuple(y.shape),
          "|d_proj_weight|", float(fp.projection.weight.grad.abs().mean()))


OPTIM_GRAD_SCALES = (3e-2, 1e-2, 1e-4)


optim_step_grad = synthetic.optim_step_grad  # the tests rebuild the gradients with the same generator


def gen_optim_step(seed=61, steps=3):
    """ref:train_byol.py:67-71 on the shimmed tiny model with the REAL torch pieces: ``clip_grad_norm_`` +
    ``torch.optim.AdamW(model.parameters(), lr, weight_decay)`` (CPU => single-tensor path) + the reference's
    ``_update_target_network``; synthetic gradients (step 0 and 1 clip, step 2 does not); two parameters never get a
    gradient (as ``masked_spec_embed`` does not in a real run)."""
    cfg_small = WavLMConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64,
                            conv_dim=(16,) * 7, num_conv_pos_embeddings=8, num_conv_pos_embedding_groups=2,
                            feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    orig = ref_encoder_mod.AutoModel.from_pretrained
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg_small))
    try:
        torch.manual_seed(seed)
        model = RefBYOL({"model": {"name": "shim", "projection_dim": 24, "prediction_dim": 48, "ema_decay": 0.996}})
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig
    rs = np.random.RandomState(seed)
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    no_grad = {named[0][0], named[5][0]}
    lr, wd = 1e-3, 1e-2
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
    out = {"names": np.array([n for n, _ in named]), "no_grad": np.array(sorted(no_grad)),
           "lr": np.array(lr), "weight_decay": np.array(wd), "ema_decay": np.array(model.ema_decay),
           "steps": np.array(steps), "seed": np.array(seed), "grad_scales": np.array(OPTIM_GRAD_SCALES)}
    pairs = list(zip(model.online_encoder.parameters(), model.target_encoder.parameters())) + \
        list(zip(model.online_projector.parameters(), model.target_projector.parameters()))
    twin = {id(o): t for o, t in pairs}
    with torch.no_grad():
        for o, t in pairs:  # make the targets differ from the online weights
            t.add_(torch.from_numpy((0.02 * rs.standard_normal(tuple(t.shape))).astype(np.float32)))
    for i, (n, p) in enumerate(named):
        out[f"p0_{i}"] = p.detach().numpy().copy()
        out[f"has_twin_{i}"] = np.array(id(p) in twin)
        if id(p) in twin:
            out[f"t0_{i}"] = twin[id(p)].detach().numpy().copy()
    norms = []
    for k, scale in zip(range(steps), OPTIM_GRAD_SCALES):
        opt.zero_grad()
        for i, (n, p) in enumerate(named):
            if n in no_grad:
                continue
            g = optim_step_grad(seed, k, i, tuple(p.shape), scale)  # regenerated by the tests, not stored
            p.grad = torch.from_numpy(g.copy())
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)))   # ref:train_byol.py:67
        opt.step()                                                                               # :70
        model._update_target_network()                                                           # :71
    out["norms"] = np.array(norms, dtype=np.float32)
    for i, (n, p) in enumerate(named):
        out[f"p_{i}"] = p.detach().numpy().copy()
        if id(p) in twin:
            out[f"t_{i}"] = twin[id(p)].detach().numpy().copy()
        if n not in no_grad:
            st = opt.state[p]
            out[f"m_{i}"], out[f"v_{i}"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "optim_step.npz"), **out)
    print("optim_step: tensors", len(named), "elems", sum(p.numel() for _, p in named), "norms", norms)


def gen_emotion(seed=71, B=5, T=37, D=64):
    """The reference's ``AttentiveStatisticsPooling`` (outputs + autograd gradients, ragged masks incl. one longer than
    T frames), ``ccc_loss`` (value + gradient) and the ``EmotionClassifier`` state-dict keys / a forward on the tiny
    shimmed encoder.  ``src.train.dimentional_emotions`` imports wandb / matplotlib at module level: stubbed."""
    import json
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from src.models.pool import AttentiveStatisticsPooling as RefASP
    from src.models.emotion import EmotionClassifier as RefEmotion
    from src.train.dimentional_emotions import ccc_loss as ref_ccc_loss
    torch.manual_seed(seed)
    rs = np.random.RandomState(seed)
    pool = RefASP(D)
    xs = torch.from_numpy(rs.standard_normal((B, T, D)).astype(np.float32)).requires_grad_(True)
    wav_lens = [T * 320 + 400, 12 * 320 + 1, 1, 20 * 320, 30 * 320 + 77]   # frames: >T (clipped), 13, 1, 20, 31
    L = max(wav_lens)
    mask = torch.zeros(B, L)
    for b, n in enumerate(wav_lens):
        mask[b, :n] = 1
    out = pool(xs, mask)
    gout = torch.from_numpy(rs.standard_normal(tuple(out.shape)).astype(np.float32))
    out.backward(gout)
    res = {"xs": xs.detach().numpy(), "mask_lens": np.array(wav_lens), "sap_w": pool.sap_linear.weight.detach().numpy(),
           "sap_b": pool.sap_linear.bias.detach().numpy(), "attention": pool.attention.detach().numpy(),
           "out": out.detach().numpy(), "gout": gout.numpy(), "dxs": xs.grad.numpy(),
           "d_sap_w": pool.sap_linear.weight.grad.numpy(), "d_sap_b": pool.sap_linear.bias.grad.numpy(),
           "d_attention": pool.attention.grad.numpy(),
           "feat_lens": np.array(pool.compute_length_from_mask(mask))}
    # a low-variance channel exercises the clamp(min=1e-5) branch
    xs2 = xs.detach().clone()
    xs2[:, :, 3] = 0.25
    xs2[:, :, 7] = 1e-3 * xs2[:, :, 7]
    xs2.requires_grad_(True)
    pool.zero_grad()
    out2 = pool(xs2, mask)
    out2.backward(gout)
    res.update({"out_clamped": out2.detach().numpy(), "dxs_clamped": xs2.grad.numpy()})
    pred = torch.from_numpy(rs.standard_normal((16, 3)).astype(np.float32)).requires_grad_(True)
    targ = torch.from_numpy((4 + 1.5 * rs.standard_normal((16, 3))).astype(np.float32))
    loss = ref_ccc_loss(pred, targ)
    loss.backward()
    res.update({"ccc_pred": pred.detach().numpy(), "ccc_targ": targ.numpy(), "ccc_loss": loss.detach().numpy(),
                "ccc_grad": pred.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "emotion.npz"), **res)
    # state-dict contract of the classifier (tiny shimmed encoder)
    cfg = small_wavlm_config("layer")
    orig = ref_encoder_mod.AutoModel.from_pretrained
    ref_encoder_mod.AutoModel.from_pretrained = staticmethod(lambda name: WavLMModel(cfg))
    try:
        enc = ref_encoder_mod.WavLMEncoder("shim")
    finally:
        ref_encoder_mod.AutoModel.from_pretrained = orig
    model = RefEmotion(enc, hidden_dim=48, dropout=0.3, num_emotions=8)
    keys = {k: list(v.shape) for k, v in model.state_dict().items() if not k.startswith("encoder.")}
    model.unfreeze_encoder_gradually([0, 1])
    trainable = sorted(n for n, p in model.encoder.model.named_parameters() if p.requires_grad)
    with open(os.path.join(HERE, "emotion_state_dict_keys.json"), "w") as f:
        json.dump({"head_keys": keys, "unfreeze_0_1": trainable}, f, indent=0, sort_keys=True)
    print("emotion: out", out.shape, "ccc_loss", float(loss), "head keys", len(keys), "unfrozen", len(trainable))


if __name__ == "__main__":
    print("transformers", transformers.__version__, "torch", torch.__version__, "numpy", np.__version__)
    gen_mix_byol()
    gen_mix_emotion()
    gen_mix_edge()
    gen_byol_loss()
    gen_ema()
    gen_frontend("layer")
    gen_frontend("group")
    gen_byol_step()
    gen_optim_step()
    gen_emotion()
    gen_evaluate_byol()
    gen_feature_projection()
    print("sizes:", {f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if f.endswith(".npz")})
