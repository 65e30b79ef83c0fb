"""Positional convolution embedding (grouped Conv1d k=128 / 16 groups with weight-norm + GELU,
hf:models/wavlm/modeling_wavlm.py:48-90): the B200 module against stock HF on the same parameters, forward and backward."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.models import B200PositionalConvEmbedding, wavlm_large_config

pytestmark = pytest.mark.gpu


def _pair(dev, seed=0):
    from transformers.models.wavlm.modeling_wavlm import WavLMPositionalConvEmbedding
    torch.manual_seed(seed)
    hf = WavLMPositionalConvEmbedding(wavlm_large_config()).to(dev)
    with torch.no_grad():   # make g and the bias non-trivial
        hf.conv.parametrizations.weight.original0.mul_(1.0 + 0.2 * torch.rand_like(hf.conv.parametrizations.weight.original0))
        hf.conv.bias.add_(0.05 * torch.randn_like(hf.conv.bias))
    mine = WavLMPositionalConvEmbedding(wavlm_large_config()).to(dev)
    mine.load_state_dict(hf.state_dict())
    assert B200PositionalConvEmbedding.supports(mine)
    return hf, B200PositionalConvEmbedding.convert(mine)


def test_pack_matches_weight_norm(dev):
    hf, mine = _pair(dev)
    w = hf.conv.weight.detach()                      # [1024, 64, 128] = g v / ||v||_tap
    wf, wb, normsq = mine._packs()
    v = hf.conv.parametrizations.weight.original1.detach()
    assert rel_err(normsq.cpu().numpy(), (v.double() ** 2).sum((0, 1)).cpu().numpy()) < 1e-5
    wf = wf.view(16, 128, 64, 64).float()            # [group, tap, n, k]
    want_f = w.view(16, 64, 64, 128).permute(0, 3, 1, 2)
    assert rel_err(wf.cpu().numpy(), want_f.cpu().numpy()) < 5e-3            # bf16 rounding
    wb = wb.view(16, 128, 64, 64).float()            # [group, tap', k, n] = w[n, k, 127 - tap']
    want_b = w.view(16, 64, 64, 128).flip(3).permute(0, 3, 2, 1)
    assert rel_err(wb.cpu().numpy(), want_b.cpu().numpy()) < 5e-3


@pytest.mark.parametrize("B,T", [(2, 24), (3, 199), (1, 130), (5, 128), (64, 199), (2, 599)])
def test_forward_matches_hf(dev, B, T):
    hf, mine = _pair(dev)
    torch.manual_seed(B * 100 + T)
    x = torch.randn(B, T, 1024, device=dev)
    with torch.no_grad():
        want = hf(x)
        got = mine(x)
    assert got.shape == want.shape == (B, T, 1024) and got.dtype == torch.float32
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-2              # bf16 operands, fp32 accumulate over K = 8192


@pytest.mark.parametrize("B,T", [(2, 24), (4, 199), (1, 130)])
def test_backward_matches_hf(dev, B, T):
    hf, mine = _pair(dev, seed=3)
    torch.manual_seed(11)
    x = torch.randn(B, T, 1024, device=dev)
    gy = torch.randn(B, T, 1024, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    hf(xa).backward(gy)
    mine(xb).backward(gy)
    assert rel_err(xb.grad.cpu().numpy(), xa.grad.cpu().numpy()) < 1.5e-2
    for (name, pa), (_, pb) in zip(hf.named_parameters(), mine.named_parameters()):
        assert pb.grad is not None and pb.grad.shape == pa.grad.shape, name
        assert rel_err(pb.grad.cpu().numpy(), pa.grad.cpu().numpy()) < 1.5e-2, name
    # frozen module, trainable input (the emotion fine-tune's gradual unfreeze never matches pos_conv_embed's names)
    for p in mine.parameters():
        p.requires_grad = False
        p.grad = None
    xc = x.clone().requires_grad_(True)
    mine(xc).backward(gy)
    assert all(p.grad is None for p in mine.parameters())
    assert rel_err(xc.grad.cpu().numpy(), xa.grad.cpu().numpy()) < 1.5e-2


def test_packs_follow_the_parameters(dev):
    hf, mine = _pair(dev)
    wf0 = mine._packs()[0].clone()
    with torch.no_grad():
        mine.conv.parametrizations.weight.original0.mul_(2.0)          # torch-side write: _version
    wf1 = mine._packs()[0]
    assert rel_err(wf1.float().cpu().numpy(), 2.0 * wf0.float().cpu().numpy()) < 1e-2
    with torch.no_grad():
        mine.conv.parametrizations.weight.original0.data.mul_(0.5)
    ops.bump_param_generation()                                        # raw-pointer writers (fused optimizer / EMA)
    assert rel_err(mine._packs()[0].float().cpu().numpy(), wf0.float().cpu().numpy()) < 1e-2


def test_installed_in_wavlm_large_encoder(dev):
    """install_b200_frontend swaps all three modules of the wavlm-large geometry; the encoder's output still matches the
    stock HF model on the same weights (tiny transformer)."""
    from transformers import WavLMModel
    from nrse_b200.models import B200FeatureEncoder, B200FeatureProjection, WavLMEncoder
    cfg = wavlm_large_config(num_hidden_layers=1, intermediate_size=64, hidden_dropout=0.0, activation_dropout=0.0,
                             attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0,
                             mask_time_prob=0.0, mask_feature_prob=0.0, apply_spec_augment=False)
    torch.manual_seed(5)
    hf = WavLMModel(cfg).to(dev).eval()
    torch.manual_seed(5)
    enc = WavLMEncoder(cfg).to(dev).eval()
    assert isinstance(enc.model.feature_extractor, B200FeatureEncoder)
    assert isinstance(enc.model.feature_projection, B200FeatureProjection)
    assert isinstance(enc.model.encoder.pos_conv_embed, B200PositionalConvEmbedding)
    x = torch.randn(2, 8000, device=dev)
    with torch.no_grad():
        want = hf(x).last_hidden_state
        got = enc(x)
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1.5e-2
