"""Feature projection (LayerNorm(512) + Linear(512 -> 1024) behind the conv stack,
hf:models/wavlm/modeling_wavlm.py:93-105): the B200 module against stock HF on the same weights, and against the fixture
produced through the reference's ``WavLMEncoder`` (tests/golden/make_golden.py::gen_feature_projection)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from nrse_b200 import ops
from nrse_b200.models import B200FeatureEncoder, B200FeatureProjection, WavLMEncoder

pytestmark = pytest.mark.gpu


def _hf_pair(dev, seed=0):
    from transformers import WavLMConfig
    from transformers.models.wavlm.modeling_wavlm import WavLMFeatureProjection
    cfg = WavLMConfig(hidden_size=1024, feat_proj_dropout=0.0)
    torch.manual_seed(seed)
    hf = WavLMFeatureProjection(cfg).to(dev)
    with torch.no_grad():
        hf.layer_norm.weight.add_(0.1 * torch.randn_like(hf.layer_norm.weight))
        hf.layer_norm.bias.add_(0.1 * torch.randn_like(hf.layer_norm.bias))
        hf.projection.bias.add_(0.1 * torch.randn_like(hf.projection.bias))
    mine = B200FeatureProjection(cfg).to(dev)
    mine.load_state_dict(hf.state_dict())
    return hf, mine


@pytest.mark.parametrize("B,T,P", [(3, 24, 24), (2, 199, 200), (64, 199, 200), (1, 130, 192)])
def test_forward_matches_hf(dev, B, T, P):
    """Inference forward on a PITCHED channels-last view (what the conv frontend hands out), fp32 and bf16 features."""
    hf, mine = _hf_pair(dev)
    torch.manual_seed(B * 1000 + T)
    buf = torch.randn(B, P, 512, device=dev) * 0.7 + 0.1
    feats = buf[:, :T]
    with torch.no_grad():
        want_h, want_n = hf(feats)
        got_h, got_n = mine(feats)
        got_hb, _ = mine(feats.bfloat16())
    assert got_h.shape == (B, T, 1024) and got_n.shape == (B, T, 512) and got_h.dtype == torch.float32
    assert rel_err(got_n.cpu().numpy(), want_n.cpu().numpy()) < 1e-5          # fp32 LayerNorm
    assert rel_err(got_h.cpu().numpy(), want_h.cpu().numpy()) < 1e-2          # bf16 GEMM operands, fp32 accumulate
    assert rel_err(got_hb.cpu().numpy(), want_h.cpu().numpy()) < 1.5e-2


@pytest.mark.parametrize("B,T", [(3, 24), (16, 199)])
def test_backward_matches_hf(dev, B, T):
    hf, mine = _hf_pair(dev, seed=1)
    torch.manual_seed(7)
    x = (torch.randn(B, T, 512, device=dev) * 0.7 + 0.1)
    g = torch.randn(B, T, 1024, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    hf.train(); mine.train()
    hf(xa)[0].backward(g)
    mine(xb)[0].backward(g)
    assert rel_err(xb.grad.cpu().numpy(), xa.grad.cpu().numpy()) < 1.5e-2
    for (name, pa), (_, pb) in zip(hf.named_parameters(), mine.named_parameters()):
        assert pb.grad is not None and rel_err(pb.grad.cpu().numpy(), pa.grad.cpu().numpy()) < 1e-2, name
    # frozen projection, trainable input (partially unfrozen encoder): only d_feats is produced
    for p in mine.parameters():
        p.requires_grad = False
        p.grad = None
    xc = x.clone().requires_grad_(True)
    mine(xc)[0].backward(g)
    assert all(p.grad is None for p in mine.parameters())
    assert rel_err(xc.grad.cpu().numpy(), xa.grad.cpu().numpy()) < 1.5e-2
    with pytest.raises(NotImplementedError):   # no gradient through the second output
        for p in mine.parameters():
            p.requires_grad = True
        mine(x.clone().requires_grad_(True))[1].sum().backward()


def test_packs_follow_the_weights(dev):
    _, mine = _hf_pair(dev)
    w16, wt16 = mine._packs()
    assert torch.equal(w16, mine.projection.weight.detach().bfloat16()) and torch.equal(wt16, w16.t())
    with torch.no_grad():
        mine.projection.weight.mul_(1.5)                    # torch-side write: _version
    assert torch.equal(mine._packs()[0], mine.projection.weight.detach().bfloat16())
    with torch.no_grad():
        mine.projection.weight.data.mul_(0.5)
    ops.bump_param_generation()                             # what the raw-pointer optimizer / EMA kernels do
    assert torch.equal(mine._packs()[0], mine.projection.weight.detach().bfloat16())


def test_golden_through_reference_encoder(dev, golden):
    """The whole chain conv frontend -> feature projection -> (stock) transformer layer inside this repository's
    ``WavLMEncoder`` against the reference's ``WavLMEncoder`` on the same seeded weights: feature-projection outputs,
    last_hidden_state, and the gradients that flow back through the projection into the conv stack."""
    from test_oracle_golden import featproj_wavlm_config
    g = golden("feature_projection")
    torch.manual_seed(int(g["seed"]))
    enc = WavLMEncoder(featproj_wavlm_config()).to(dev).train()
    assert isinstance(enc.model.feature_projection, B200FeatureProjection)
    assert isinstance(enc.model.feature_extractor, B200FeatureEncoder)
    captured = {}
    h = enc.model.feature_projection.register_forward_hook(lambda m, i, o: captured.update(hidden=o[0], norm=o[1]))
    y = enc(torch.from_numpy(g["x"]).to(dev)[:, None])
    h.remove()
    assert rel_err(captured["norm"].detach().cpu().numpy(), g["norm"]) < 1e-2
    assert rel_err(captured["hidden"].detach().cpu().numpy(), g["hidden"]) < 1e-2
    assert rel_err(y.detach().cpu().numpy(), g["last_hidden"]) < 1e-2
    (y * torch.from_numpy(g["G"]).to(dev)).sum().backward()
    fp = enc.model.feature_projection
    fe = enc.model.feature_extractor
    errs = {
        "d_ln_weight": rel_err(fp.layer_norm.weight.grad.cpu().numpy(), g["d_ln_weight"]),
        "d_ln_bias": rel_err(fp.layer_norm.bias.grad.cpu().numpy(), g["d_ln_bias"]),
        "d_proj_bias": rel_err(fp.projection.bias.grad.cpu().numpy(), g["d_proj_bias"]),
        "d_proj_weight": rel_err(fp.projection.weight.grad.cpu().numpy()[::64], g["d_proj_weight_rows"]),
        "d_conv6_weight": rel_err(fe.conv_layers[6].conv.weight.grad.cpu().numpy()[::64], g["d_conv6_weight_rows"]),
        "d_conv0_weight": rel_err(fe.conv_layers[0].conv.weight.grad.cpu().numpy(), g["d_conv0_weight"]),
    }
    print("feature projection golden:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 2e-2, errs
