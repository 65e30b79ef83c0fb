"""CPU tests of the host-side mirror of the reference surface: config, logging, dataset I/O, module/state-dict
contract, and the data-parallel plumbing over a 2-rank gloo group.  No kernel is launched here."""
import json
import os
import wave

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from nrse_b200.config import get_config, load_config
from nrse_b200.data import MixedBatchLoader, NoiseRobustSpeechDataset, TensorPairDataset, create_dataloaders
from nrse_b200.data.noisy_speech_dataset import load_and_process_audio
from nrse_b200.models import B200FeatureEncoder, BYOLSpeechModel, WavLMEncoder, wavlm_large_config
from nrse_b200.train import EarlyStopping
from nrse_b200.utils.setup_utils import set_seed


def small_config(norm="layer"):
    from transformers import WavLMConfig
    return WavLMConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                       feat_extract_norm=norm, do_stable_layer_norm=(norm == "layer"), conv_bias=False,
                       num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)


def golden_config():
    """The configuration tests/golden/make_golden.py::gen_byol_step used: deterministic in train mode."""
    cfg = small_config("layer")
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout", "feat_proj_dropout", "final_dropout",
              "layerdrop", "mask_time_prob", "mask_feature_prob"):
        setattr(cfg, k, 0.0)
    cfg.apply_spec_augment = False
    return cfg


def byol_config(wavlm=None):
    return {"model": {"name": wavlm or golden_config(), "projection_dim": 96, "prediction_dim": 128, "ema_decay": 0.99}}


def test_state_dict_contract_matches_reference():
    """Keys and shapes of BYOLSpeechModel.state_dict() are those of the reference's model (fixture written by
    tests/golden/make_golden.py from the unmodified reference): checkpoints round-trip both ways."""
    want = json.load(open(os.path.join(GOLDEN, "byol_state_dict_keys.json")))
    model = BYOLSpeechModel(byol_config())
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    assert isinstance(model.online_encoder.model.feature_extractor, B200FeatureEncoder)
    assert all(not p.requires_grad for p in model.target_encoder.parameters())
    assert all(not p.requires_grad for p in model.target_projector.parameters())
    assert all(p.requires_grad for p in model.online_predictor.parameters())
    for o, t in zip(model.online_encoder.parameters(), model.target_encoder.parameters()):
        assert torch.equal(o, t)  # target starts as a copy of the online weights
    assert model.get_encoder() is model.online_encoder
    # the emotion fine-tune loop unfreezes by substring match on parameter names (ref:src/models/emotion.py:126-129)
    names = [n for n, _ in model.online_encoder.model.named_parameters()]
    assert any("conv_layers.3" in n and "layers.3" in n for n in names)


def test_frontend_conversion_keeps_parameters_and_modes():
    from transformers.models.wavlm.modeling_wavlm import WavLMFeatureEncoder
    for norm in ("layer", "group"):
        fe = WavLMFeatureEncoder(small_config(norm))
        before = {k: v.clone() for k, v in fe.state_dict().items()}
        conv = B200FeatureEncoder.convert(fe)
        assert conv is fe and isinstance(fe, B200FeatureEncoder) and fe.norm_mode == norm
        assert list(fe.state_dict().keys()) == list(before.keys())
        assert all(torch.equal(fe.state_dict()[k], v) for k, v in before.items())
        fe._freeze_parameters()
        assert not any(p.requires_grad for p in fe.parameters()) and fe._requires_grad is False
    from transformers import WavLMConfig
    with pytest.raises(ValueError):
        B200FeatureEncoder.convert(WavLMFeatureEncoder(WavLMConfig(conv_bias=True)))
    cfg = wavlm_large_config()
    assert cfg.hidden_size == 1024 and cfg.feat_extract_norm == "layer" and tuple(cfg.conv_kernel) == (10, 3, 3, 3, 3, 2, 2)


def test_encoder_wrapper_surface_cpu():
    """frontend='hf' keeps the stock feature extractor, so the wrapper can be exercised without a GPU: it must behave
    like the reference's WavLMEncoder (squeeze [B,1,L], ignore the mask, return last_hidden_state)."""
    torch.manual_seed(0)
    enc = WavLMEncoder(small_config(), frontend="hf").eval()
    x = torch.randn(2, 1, 2000)
    with torch.no_grad():
        a = enc(x)
        b = enc(x.squeeze(1), attention_mask=torch.ones(2, 2000))
        c = enc.model(x.squeeze(1)).last_hidden_state
    assert a.shape == (2, 6, 64) and enc.output_dim == 64
    assert torch.equal(a, b) and torch.equal(a, c)
    with pytest.raises(ValueError):
        WavLMEncoder(small_config(), frontend="nope")


def test_b200_frontend_refuses_cpu():
    from nrse_b200._lib import NrseError
    enc = WavLMEncoder(small_config(), frontend="b200").eval()
    with pytest.raises(NrseError), torch.no_grad():
        enc(torch.randn(1, 2000))


def test_config_loader_and_overrides(tmp_path):
    y = tmp_path / "c.yaml"
    y.write_text("model:\n  name: m\n  ema_decay: 0.997\ntraining:\n  batch_size: 36\n  learning_rate: 1.0e-5\n"
                 "data:\n  snr_range: [2, 5]\n")
    cfg = load_config(str(y))
    assert cfg["model"]["ema_decay"] == 0.997 and cfg["data"]["snr_range"] == [2, 5]
    cfg = get_config(["--config", str(y), "--device", "cuda:1", "--batch_size", "8", "--epochs", "3", "--lr", "0.01"])
    assert cfg["device"] == "cuda:1" and cfg["training"]["batch_size"] == 8
    assert cfg["training"]["num_epochs"] == 3 and cfg["training"]["learning_rate"] == 0.01


def test_early_stopping_and_seed():
    es = EarlyStopping(patience=2, min_delta=0.01, mode="min")
    assert [es(v) for v in (1.0, 0.9, 0.895, 0.894)] == [False, False, False, True]
    es = EarlyStopping(patience=1, mode="max")
    assert [es(v) for v in (0.5, 0.6, 0.6)] == [False, False, True]
    set_seed(3); a = (torch.rand(2), np.random.rand())
    set_seed(3); b = (torch.rand(2), np.random.rand())
    assert torch.equal(a[0], b[0]) and a[1] == b[1]


def _write_wav(path, samples, sr=16000):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes((np.clip(samples, -1, 1) * 32767).astype("<i2").tobytes())


def test_dataset_raw_items_from_wav_files(tmp_path):
    rs = np.random.RandomState(0)
    cdir, ndir = tmp_path / "clean", tmp_path / "noise"
    cdir.mkdir(); ndir.mkdir()
    _write_wav(cdir / "a.wav", 0.1 * rs.standard_normal(20000))
    _write_wav(cdir / "b.wav", 0.1 * rs.standard_normal(3000))      # shorter than max length: zero padded
    _write_wav(cdir / "silent.wav", np.zeros(8000))                  # rejected, dataset moves on to the next file
    _write_wav(ndir / "n.wav", 0.3 * rs.standard_normal(9000))
    (cdir / "notes.txt").write_text("x")
    ds = NoiseRobustSpeechDataset(str(cdir), str(ndir), 16000, 0.5, [2, 5, 10])
    assert len(ds) == 3 and ds.max_samples == 8000
    set_seed(1)
    for i in range(3):
        item = ds[i]
        assert item["clean_wave"].shape == (1, 8000) and item["noise_wave"].shape == (1, 8000)
        assert item["snr"] in (2, 5, 10) and ds.snr_range[item["snr_idx"]] == item["snr"]
        assert float(item["clean_wave"].abs().max()) > 1e-3
    short = load_and_process_audio(str(cdir / "b.wav"), 16000, 0.5)
    assert short.shape == (1, 8000) and not short[0, 3000:].any()
    assert load_and_process_audio(str(cdir / "silent.wav"), 16000, 0.5) is None
    assert load_and_process_audio(str(cdir / "missing.wav"), 16000, 0.5) is None


def test_create_dataloaders_split_is_seeded():
    clean, noise = torch.randn(20, 800), torch.randn(20, 800)
    ds = TensorPairDataset(clean, noise, [2, 5, 10, 15, 20])
    cfg = {"data": {"snr_range": [2, 5, 10, 15, 20], "validation_ratio": 0.15, "clean_data_path": "", "noise_data_path": "",
                    "sample_rate": 16000, "max_audio_length": 0.05},
           "training": {"batch_size": 4, "num_workers": 0, "seed": 42}, "device": "cpu"}
    tr, va = create_dataloaders(cfg, dataset=ds)
    tr2, va2 = create_dataloaders(cfg, dataset=ds)
    assert isinstance(tr, MixedBatchLoader) and len(tr.dataset) == 17 and len(va.dataset) == 3 and len(tr) == 5
    assert list(va.dataset.indices) == list(va2.dataset.indices)
    item = ds[7]
    assert item["snr"] == [2, 5, 10, 15, 20][7 % 5] and item["clean_wave"].shape == (1, 800)


def perturbed_eval_model(g, frontend):
    """The model of tests/golden/make_golden.py::gen_evaluate_byol: seeded construction, then the same seeded
    perturbation of the online encoder (so that online / target branches differ like in a trained model)."""
    torch.manual_seed(int(g["seed"]))
    cfg = byol_config(golden_config())
    cfg["model"]["frontend"] = frontend
    model = BYOLSpeechModel(cfg)
    with torch.no_grad():
        gen = torch.Generator().manual_seed(int(g["seed"]))
        for p in model.online_encoder.parameters():
            p.add_(0.02 * p.abs().mean() * torch.randn(p.shape, generator=gen))
    return model


def test_evaluate_oracle_matches_reference_fixture():
    """oracle.embedding_similarity / validation_metrics against the values the reference's own
    evaluate_embedding_similarity / validate_model (ref:evaluate_byol.py:12-123) produced -- through this repository's
    BYOLSpeechModel with the stock HF frontend on the CPU (same state-dict, same seeded init)."""
    g = np.load(os.path.join(GOLDEN, "evaluate_byol.npz"))
    model = perturbed_eval_model(g, "hf").eval()
    snr_range = g["snr_range"].tolist()
    embs, pairs = [], []
    with torch.no_grad():
        for i in range(3):
            c, n = torch.from_numpy(g[f"clean_{i}"])[:, None], torch.from_numpy(g[f"noisy_{i}"])[:, None]
            enc = model.get_encoder()
            embs.append((model._pool(enc(c)), model._pool(enc(n)), g[f"snr_{i}"]))
            pairs.append(model(c, n))
    sims = oracle_mod().embedding_similarity(embs, snr_range)
    val_loss, metrics = oracle_mod().validation_metrics(pairs, sims)
    want = dict(zip(snr_range, g["similarities"].tolist()))
    for s in snr_range:
        assert abs(sims[s] - want[s]) < 1e-5, (s, sims[s], want[s])
    assert sims[15] == 0 and abs(val_loss - float(g["val_loss"])) < 1e-5
    assert abs(metrics["val_avg_similarity"] - float(g["val_avg_similarity"])) < 1e-5


def oracle_mod():
    import oracle
    return oracle


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import oracle
    from nrse_b200.train import init_distributed, wrap_data_parallel
    r, w, lr = init_distributed("gloo")
    assert (r, w, lr) == (rank, world, rank)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(32, 48), torch.nn.ReLU(), torch.nn.Linear(48, 16))
    frozen = torch.nn.Linear(32, 16)  # stands for the target branch: no grad, never all-reduced
    for p in frozen.parameters():
        p.requires_grad = False
    model = torch.nn.ModuleDict({"online": net, "target": frozen})
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 32, generator=g)
    shard = x[rank * 4:(rank + 1) * 4]  # batch sharded across ranks, no data-path collective

    class Both(torch.nn.Module):  # online + (frozen) target branch behind one forward, like BYOLSpeechModel
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, v):
            return self.m["online"](v), self.m["target"](v)

    both = wrap_data_parallel(Both(model), torch.device("cpu"))
    p, z = both(shard)
    oracle.byol_loss(p, z.detach()).backward()
    grads = torch.cat([q.grad.flatten() for q in net.parameters()])
    # single-process reference on the full batch: per-rank means average to the global mean for equal shards
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(32, 48), torch.nn.ReLU(), torch.nn.Linear(48, 16))
    oracle.byol_loss(ref(x), frozen(x).detach()).backward()
    ref_grads = torch.cat([q.grad.flatten() for q in ref.parameters()])
    ok = torch.allclose(grads, ref_grads, rtol=1e-5, atol=1e-7) and all(q.grad is None for q in frozen.parameters())
    gathered = [torch.zeros_like(grads) for _ in range(world)]
    dist.all_gather(gathered, grads)
    same = all(torch.equal(gathered[0], t) for t in gathered)  # every rank holds the same averaged gradient
    out[rank] = bool(ok and same)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_equivalence():
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def _arena_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import oracle
    from nrse_b200.data import TensorPairDataset
    from nrse_b200.train import GradArena, init_distributed
    init_distributed("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(32, 48), torch.nn.ReLU(), torch.nn.Linear(48, 15))
    extra = [torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(3, 5))]  # grads written by hand
    arena = GradArena([extra, list(net.parameters())], bucket_bytes=1024)   # several buckets per group
    ptrs = [p.grad.data_ptr() for p in net.parameters()]
    assert all(p.grad.data_ptr() % 16 == 0 for g in arena.groups for p in g)
    g = torch.Generator().manual_seed(1)
    x, z = torch.randn(8, 32, generator=g), torch.randn(8, 15, generator=g)
    ok = True
    for step in range(2):
        arena.zero_(1)
        for i, p in enumerate(extra):
            p.grad.fill_(float(rank + 1 + i))                       # group 0: not from autograd
        n0 = arena.all_reduce_async(0)                               # travels while the "backward" below runs
        oracle.byol_loss(net(x[rank * 4:(rank + 1) * 4]), z[rank * 4:(rank + 1) * 4]).backward()   # accumulates into the views
        n1 = arena.all_reduce_async(1)
        arena.wait()
        ok &= n0 >= 1 and n1 >= 2 and [p.grad.data_ptr() for p in net.parameters()] == ptrs
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(32, 48), torch.nn.ReLU(), torch.nn.Linear(48, 15))
        oracle.byol_loss(ref(x), z).backward()
        for a, b in zip(net.parameters(), ref.parameters()):
            ok &= torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7)
        for i, p in enumerate(extra):
            ok &= torch.allclose(p.grad, torch.full_like(p, (1 + 2) / 2 + i))
    # create_dataloaders shards the dataset across ranks (DistributedSampler) once a process group exists
    from nrse_b200.data import noisy_speech_dataset as nsd
    class _NoMixer:  # the GPU mixer is not under test here
        def __init__(self, *a, **k): pass
        def __call__(self, raw): return raw
    nsd.GpuBatchMixer = _NoMixer
    ds = TensorPairDataset(torch.arange(40.0)[:, None].repeat(1, 8), torch.zeros(40, 8), [2, 5])
    cfg = {"data": {"snr_range": [2, 5], "validation_ratio": 0.2}, "training": {"batch_size": 4, "num_workers": 0, "seed": 3}}
    tr, va = nsd.create_dataloaders(cfg, device="cpu", dataset=ds)
    tr.set_epoch(0)
    mine = sorted(int(v) for b in tr for v in b["clean_wave"][:, 0, 0].tolist())
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok &= len(mine) == 16 and not set(gathered[0]) & set(gathered[1]) and len(set(gathered[0]) | set(gathered[1])) == 32
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_grad_arena_and_sharded_loader():
    """GradArena: gradients live in one flat buffer (stable addresses), groups are all-reduced asynchronously in several
    buckets, the mean over ranks equals the full-batch gradient; hand-written gradients take part.  create_dataloaders
    hands every rank a disjoint shard."""
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_arena_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_emotion_classifier_contract_matches_reference():
    """Head state-dict keys / shapes and the gradual-unfreeze substring match equal the reference's
    (tests/golden/emotion_state_dict_keys.json, written from ref:src/models/emotion.py on the shimmed encoder)."""
    from nrse_b200.models import EmotionClassifier
    with open(os.path.join(GOLDEN, "emotion_state_dict_keys.json")) as f:
        want = json.load(f)
    enc = WavLMEncoder(golden_config())
    model = EmotionClassifier(enc, hidden_dim=48, dropout=0.3, num_emotions=8)
    got = {k: list(v.shape) for k, v in model.state_dict().items() if not k.startswith("encoder.")}
    assert got == want["head_keys"]
    model.unfreeze_encoder_gradually([0, 1])
    trainable = sorted(n for n, p in model.encoder.model.named_parameters() if p.requires_grad)
    assert trainable == want["unfreeze_0_1"]
    model.freeze_encoder()
    assert all(not p.requires_grad for p in model.encoder.parameters())
    n_head = model.get_trainable_params()
    model.unfreeze_encoder()
    assert model.get_trainable_params() > n_head
    # compute_length_from_mask: same numbers as the reference's list, but a device tensor
    lens = model.pooling.compute_length_from_mask(torch.tensor([[1.0] * 640, [1.0] * 321 + [0.0] * 319]))
    assert lens.tolist() == [2, 2] and lens.dtype == torch.int32


def test_fused_optimizer_host_logic_without_gpu(monkeypatch):
    """FusedAdamWEma's bookkeeping (bucketing by step count, lazy step counters, state_dict interchange with
    torch.optim.AdamW, EMA twins of gradient-less parameters) with the device table replaced by a recorder: no kernel
    runs, so this covers the host side on the CPU; the arithmetic is covered by tests/test_gpu_optim.py."""
    from nrse_b200.train import optim as O
    calls = []

    class Recorder:
        @staticmethod
        def partials_count():
            return 4

        def update(self, params, grads, m, v, twins):
            calls.append(("update", [g is not None for g in grads], [t is not None for t in twins]))

        def grad_sqnorm(self, partials):
            calls.append(("norm", partials.numel()))

        def clip_adamw_ema(self, **kw):
            calls.append(("step", kw["step"], kw["max_grad_norm"], kw["ema_decay"], kw["lr"]))

    monkeypatch.setattr(O.ops, "OptimChunkTable", Recorder)
    P = [torch.nn.Parameter(torch.randn(5)) for _ in range(3)]
    twin = torch.randn(5)
    opt = O.FusedAdamWEma(P, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0, ema_pairs=[(P[2], twin)], ema_decay=0.99)
    assert opt.has_ema and isinstance(opt, torch.optim.AdamW)
    with pytest.raises(ValueError):
        opt.attach_ema([(torch.nn.Parameter(torch.randn(5)), twin)], 0.9)  # not owned by this optimizer
    opt.attach_ema([(P[2], twin)], 0.99)
    grads = [torch.randn(5) for _ in range(3)]
    for k in range(3):
        for i, p in enumerate(P):
            p.grad = None if (k == 0 and i == 2) else grads[i]  # P[2] joins one step late; its twin is an orphan first
        opt.step()
    steps = [c for c in calls if c[0] == "step"]
    assert [c[1] for c in steps] == [1, 1, 2, 2, 3]        # k=0: one bucket; k=1, k=2: P[2] one step behind
    assert calls[0] == ("update", [True, True, False], [False, False, True])
    assert all(c[2] == 1.0 and c[3] == 0.99 for c in steps)
    assert opt.table_builds == 2                            # k=0 and k=1 (new gradient pattern); k=2 reuses the tables
    sd = opt.state_dict()
    assert {k: float(v["step"]) for k, v in sd["state"].items()} == {0: 3.0, 1: 3.0, 2: 2.0}
    # a torch.optim.AdamW resumes from it, and vice versa
    ref = torch.optim.AdamW(P, lr=1e-3, weight_decay=1e-2)
    ref.load_state_dict(sd)
    assert float(ref.state[P[2]]["step"]) == 2.0
    opt2 = O.FusedAdamWEma(P, lr=1e-3, weight_decay=1e-2)
    opt2.load_state_dict(ref.state_dict())
    calls.clear()
    opt2.step()
    assert sorted(c[1] for c in calls if c[0] == "step") == [3, 4] and not any(c[0] == "norm" for c in calls)
    # LR schedulers act on param_groups as usual
    sched = torch.optim.lr_scheduler.StepLR(opt2, step_size=1, gamma=0.5)
    sched.step()
    calls.clear()
    opt2.step()
    assert all(abs(c[4] - 5e-4) < 1e-12 for c in calls if c[0] == "step")
    # zero_grad follows torch's default unless keep_grads is set
    opt2.zero_grad()
    assert all(p.grad is None for p in P)


def test_sync_free_spec_augment_equals_stock_masking():
    """The sync-free SpecAugment replacement produces the same hidden states and gradients as HF's
    ``_mask_hidden_states`` for the same numpy RNG state (drawn time masks and explicit mask_time_indices)."""
    from transformers import WavLMModel
    from nrse_b200.models import install_sync_free_spec_augment
    cfg = small_config("layer")
    cfg.mask_time_prob, cfg.mask_time_length = 0.3, 2  # (feature masking: WavLMConfig lacks mask_feature_min_masks,
    # the stock method raises there; wavlm-large has mask_feature_prob = 0)
    torch.manual_seed(0)
    stock = WavLMModel(cfg).train()
    mine = WavLMModel(cfg).train()
    mine.load_state_dict(stock.state_dict())
    install_sync_free_spec_augment(mine)
    h = torch.randn(3, 24, cfg.hidden_size)
    for explicit in (None, torch.rand(3, 24) < 0.25):
        a = h.clone().requires_grad_(True)
        b = h.clone().requires_grad_(True)
        np.random.seed(5)
        want = stock._mask_hidden_states(a * 1.0, mask_time_indices=explicit)
        np.random.seed(5)
        got = mine._mask_hidden_states(b * 1.0, mask_time_indices=explicit)
        assert torch.equal(got, want)
        (want * want).sum().backward()
        (got * got).sum().backward()
        assert torch.equal(a.grad, b.grad)
        assert mine.masked_spec_embed.grad is not None
    mine.eval()
    assert torch.equal(mine._mask_hidden_states(h.clone()), h)  # inference: no masking
