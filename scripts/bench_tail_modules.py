"""Device time of the two modules behind the conv stack -- feature projection (8f-1) and positional conv (8f-4) -- on the
B200 kernels and as stock HF modules on the same GPU, at the BASELINE shape (64 utterances x 4 s -> 64 x 199 frames)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers.models.wavlm.modeling_wavlm import WavLMFeatureProjection, WavLMPositionalConvEmbedding
from nrse_b200.models import B200FeatureProjection, B200PositionalConvEmbedding, wavlm_large_config

dev = torch.device("cuda:0")
cfg = wavlm_large_config(feat_proj_dropout=0.0)
B, T = 64, 199


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"shape": [B, T]}
for name, hf_cls, mine_cls, cin, autocast in (("feature_projection", WavLMFeatureProjection, B200FeatureProjection, 512, True),
                                              ("pos_conv_embed", WavLMPositionalConvEmbedding, B200PositionalConvEmbedding, 1024, True)):
    torch.manual_seed(0)
    hf = hf_cls(cfg).to(dev).train()
    mine = hf_cls(cfg).to(dev).train()
    mine.load_state_dict(hf.state_dict())
    mine = mine_cls.convert(mine)
    x = torch.randn(B, T, cin, device=dev)
    gy = torch.randn(B, T, 1024, device=dev)
    out = (lambda m, v: m(v)[0]) if name == "feature_projection" else (lambda m, v: m(v))

    def fwd(m, ac=False):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            return out(m, x)

    def fwd_bwd(m, ac=False):
        xx = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            y = out(m, xx)
        y.backward(gy.to(y.dtype))

    r = {"b200_fwd_ms": timeit(lambda: fwd(mine)), "b200_fwd_bwd_ms": timeit(lambda: fwd_bwd(mine)),
         "hf_fp32_fwd_ms": timeit(lambda: fwd(hf)), "hf_fp32_fwd_bwd_ms": timeit(lambda: fwd_bwd(hf)),
         "hf_bf16_autocast_fwd_ms": timeit(lambda: fwd(hf, True)), "hf_bf16_autocast_fwd_bwd_ms": timeit(lambda: fwd_bwd(hf, True))}
    if name == "pos_conv_embed":
        flops = 2.0 * B * T * 1024 * 64 * 128
        r["algorithmic_gflop_fwd"] = flops / 1e9
        r["b200_fwd_tflops"] = flops / (r["b200_fwd_ms"] * 1e-3) / 1e12
    else:
        flops = 2.0 * B * T * 512 * 1024
        r["algorithmic_gflop_fwd"] = flops / 1e9
    res[name] = r
print(json.dumps(res))
