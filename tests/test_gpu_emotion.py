"""GPU parity of the emotion fine-tune consumer (ref:src/models/pool.py, ref:src/models/emotion.py,
ref:src/train/dimentional_emotions.py): batched attentive statistics pooling forward/backward against the fixtures
written by the reference's own classes and against the oracle at ragged sizes, the classifier on the B200 encoder, and
one fine-tune step."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.models import AttentiveStatisticsPooling, EmotionClassifier, WavLMEncoder
from nrse_b200.train import FusedAdamWEma, ccc_loss, emotion_dim_step, train_one_epoch_dimensional
from test_host_modules import golden_config
from test_oracle_golden import _emotion_mask

pytestmark = pytest.mark.gpu
TOL = 2e-5  # fp32 with different tanh / exp implementations and reduction orders (reference: CPU libm)


def _pool_from(g, dev):
    D = g["sap_w"].shape[0]
    pool = AttentiveStatisticsPooling(D).to(dev)
    with torch.no_grad():
        pool.sap_linear.weight.copy_(torch.from_numpy(g["sap_w"]))
        pool.sap_linear.bias.copy_(torch.from_numpy(g["sap_b"]))
        pool.attention.copy_(torch.from_numpy(g["attention"]))
    return pool


def test_asp_golden_forward_backward(dev, golden):
    g = golden("emotion")
    pool = _pool_from(g, dev)
    mask = _emotion_mask(g).to(dev)
    for clamped in (False, True):
        xs = torch.from_numpy(g["xs"].copy())
        if clamped:
            xs[:, :, 3] = 0.25
            xs[:, :, 7] = 1e-3 * xs[:, :, 7]
        xs = xs.to(dev).requires_grad_(True)
        pool.zero_grad()
        out = pool(xs, mask)
        out.backward(torch.from_numpy(g["gout"]).to(dev))
        assert rel_err(out.detach().cpu().numpy(), g["out_clamped" if clamped else "out"]) < TOL
        assert rel_err(xs.grad.cpu().numpy(), g["dxs_clamped" if clamped else "dxs"]) < TOL
        if not clamped:
            assert rel_err(pool.sap_linear.weight.grad.cpu().numpy(), g["d_sap_w"]) < TOL
            assert rel_err(pool.sap_linear.bias.grad.cpu().numpy(), g["d_sap_b"]) < TOL
            assert rel_err(pool.attention.grad.cpu().numpy(), g["d_attention"]) < TOL
    # padded frames receive exactly zero gradient (the reference's slice never touches them)
    lens = [int(v) for v in g["feat_lens"]]
    for b, n in enumerate(lens):
        assert not xs.grad[b, min(n, xs.shape[1]):].any()


@pytest.mark.parametrize("B,T,D,att_scale,tol", [(3, 199, 1024, 0.03, TOL), (1, 1, 64, 1.0, TOL), (7, 50, 256, 0.06, TOL),
                                                 (2, 600, 1024, 0.03, TOL), (4, 33, 36, 0.2, TOL),
                                                 (3, 199, 1024, 1.0, 1e-3)])
def test_asp_ragged_vs_oracle(dev, B, T, D, att_scale, tol):
    """``att_scale`` ~ 1/sqrt(D) keeps the attention logits O(1).  The last case uses the reference's N(0,1) init at
    D = 1024: logits of +-60, a softmax that is one-hot to 1e-20 and amplifies the 1e-5 absolute rounding noise of a
    1024-term fp32 dot product into 1e-4 relative weight differences -- on the reference's side as much as here."""
    rs = np.random.RandomState(B * 1000 + T)
    xs = torch.from_numpy(rs.standard_normal((B, T, D)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32))
    bias = torch.from_numpy(0.1 * rs.standard_normal(D).astype(np.float32))
    att = torch.from_numpy((att_scale * rs.standard_normal((D, 1))).astype(np.float32))
    frames = [max(1, int(rs.randint(1, T + 1))) for _ in range(B)]
    frames[0] = T
    L = (T - 1) * 320 + 400
    mask = torch.zeros(B, L)
    for b, f in enumerate(frames):
        mask[b, :(f - 1) * 320 + 1] = 1
    assert [min(v, T) for v in oracle.compute_length_from_mask(mask)] == frames
    gout = torch.from_numpy(rs.standard_normal((B, 2 * D)).astype(np.float32))
    cx, cw, cb, ca = (t.clone().requires_grad_(True) for t in (xs, w, bias, att))
    want = oracle.attentive_statistics_pooling(cx, mask, cw, cb, ca)
    want.backward(gout)
    pool = AttentiveStatisticsPooling(D).to(dev)
    with torch.no_grad():
        pool.sap_linear.weight.copy_(w); pool.sap_linear.bias.copy_(bias); pool.attention.copy_(att)
    gx = xs.to(dev).requires_grad_(True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False  # the sap_linear GEMM in full fp32, like the reference
    try:
        got = pool(gx, mask.to(dev))
        got.backward(gout.to(dev))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert rel_err(got.detach().cpu().numpy(), want.detach().numpy()) < tol
    assert rel_err(gx.grad.cpu().numpy(), cx.grad.numpy()) < 5 * tol
    assert rel_err(pool.attention.grad.cpu().numpy(), ca.grad.numpy()) < 5 * tol
    assert rel_err(pool.sap_linear.weight.grad.cpu().numpy(), cw.grad.numpy()) < 5 * tol


def test_asp_rejects_cpu_and_bad_shapes(dev):
    x = torch.randn(2, 5, 8)
    with pytest.raises(Exception):
        ops.asp_pool(x, x, torch.randn(8), torch.tensor([5, 5]))
    xd = x.to(dev)
    with pytest.raises(Exception):
        ops.asp_pool(xd, xd[:, :4], torch.randn(8, device=dev), torch.tensor([5, 5], device=dev))
    with pytest.raises(Exception):
        ops.asp_pool(xd[:, :, :6].contiguous(), xd[:, :, :6].contiguous(), torch.randn(6, device=dev),
                     torch.tensor([5, 5], device=dev))


def test_emotion_classifier_forward_and_finetune_step(dev):
    """EmotionClassifier on the B200 encoder: the pooled features equal the oracle's pooling of the same encoder
    output; a dimensional fine-tune step with the fused optimizer tail moves the unfrozen parameters only."""
    torch.manual_seed(5)
    enc = WavLMEncoder(golden_config())
    model = EmotionClassifier(enc, hidden_dim=64, dropout=0.0, num_emotions=8).to(dev)
    model.eval()
    x = torch.randn(4, 8000, device=dev)
    mask = torch.ones(4, 8000, device=dev)
    mask[1, 4000:] = 0
    mask[3, 6500:] = 0
    with torch.no_grad():
        h = model.encoder(x, attention_mask=mask)
        feats = model.pooling(h, mask)
        cat, dim = model(x, attention_mask=mask, task="both")
    want = oracle.attentive_statistics_pooling(h.float().cpu(), mask.cpu(), model.pooling.sap_linear.weight.detach().cpu(),
                                               model.pooling.sap_linear.bias.detach().cpu(),
                                               model.pooling.attention.detach().cpu())
    assert rel_err(feats.cpu().numpy(), want.numpy()) < 1e-3   # sap_linear GEMM may run in TF32 on the GPU
    assert cat.shape == (4, 8) and dim.shape == (4, 3)
    assert model(x, task="categorical")[1] is None and model(x, task="dimensional")[0] is None
    # fine-tune step: encoder frozen except transformer layer 1 + conv layer 1 (substring match), heads trainable
    model.train()
    model.unfreeze_encoder_gradually([1])
    opt = FusedAdamWEma(model.parameters(), lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    labels = torch.rand(4, 3, device=dev) * 6 + 1
    loss, values = emotion_dim_step(model, x, labels, opt, mask)
    assert torch.isfinite(loss) and values.shape == (4, 3)
    moved = {n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])}
    frozen = {n for n, p in model.named_parameters() if not p.requires_grad}
    assert moved and not (moved & frozen)
    assert any("feature_extractor.conv_layers.1." in n for n in moved)       # native frontend backward reached it
    assert any(n.startswith("pooling.") for n in moved) and any(n.startswith("dimensional_out") for n in moved)
    # epoch loop
    batches = [{"input_values": x.cpu(), "attention_mask": mask.cpu(), "A": labels[:, 0].cpu(), "V": labels[:, 1].cpu(),
                "D": labels[:, 2].cpu()} for _ in range(2)]
    mean_loss, ccc = train_one_epoch_dimensional(model, batches, opt, dev)
    assert np.isfinite(mean_loss) and set(ccc) == {"A", "V", "D", "avg"}
