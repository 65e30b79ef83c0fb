"""Drop-in replacement of HF ``WavLMFeatureProjection`` (hf:models/wavlm/modeling_wavlm.py:93-105): LayerNorm(512) +
Linear(512 -> 1024) + dropout, the step right behind the conv feature encoder in ``WavLMModel.forward``
(hf:...:1061-1064, reached from ref:src/models/encoder.py:25).  SURVEY.md 8f-1.

It IS a ``WavLMFeatureProjection`` (same ``layer_norm`` / ``projection`` / ``dropout`` sub-modules, same parameter names
and shapes: checkpoints and ``named_parameters()`` are untouched); ``forward`` runs the B200 kernels: one LayerNorm launch
that reads the conv frontend's pitched channels-last output IN PLACE and writes the bf16 GEMM operand, one tcgen05 GEMM
launch with an fp32 + bias epilogue -- no transpose, no separate LayerNorm pass over fp32, no cuBLAS call.  Backward (four
launches): gradient cast + bias gradient, tcgen05 weight-gradient GEMM, tcgen05 data-gradient GEMM, LayerNorm backward
whose fp32 output is exactly the ``dy`` the conv frontend's native backward consumes.

Only the wavlm-large geometry (512 -> 1024) is implemented; other sizes keep the stock module (``convert`` refuses them).
The gradient through the second output (``norm_hidden_states`` / ``extract_features``) is not implemented: nothing in
``WavLMModel`` or the reference uses it.
"""
from __future__ import annotations

import torch
from transformers.models.wavlm.modeling_wavlm import WavLMFeatureProjection

from .. import ops


class _FeatProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, module, ln_w, ln_b, w, b):
        w16, wt16 = module._packs()
        hidden, norm, tape = ops.feature_projection_fwd(feats, ln_w, ln_b, module.layer_norm.eps, w16, b, training=True)
        ctx.save_for_backward(ln_w, ln_b)
        ctx.tape, ctx.wt16 = tape, wt16
        ctx.set_materialize_grads(False)
        return hidden, norm

    @staticmethod
    def backward(ctx, g_hidden, g_norm):
        if g_norm is not None:
            raise NotImplementedError("B200FeatureProjection: the gradient through norm_hidden_states (extract_features) "
                                      "is not implemented -- nothing in WavLMModel or the reference uses it")
        if ctx.tape is None:
            raise RuntimeError("B200FeatureProjection: the tape of this forward was freed by its first backward")
        if g_hidden is None:
            return (None,) * 6
        ln_w, ln_b = ctx.saved_tensors
        need = ctx.needs_input_grad
        d_feats, d_g, d_b, d_w, d_bias = ops.feature_projection_bwd(
            g_hidden, ctx.tape, ln_w, ln_b, ctx.wt16, need_feats=need[0], need_ln=need[2] or need[3], need_w=need[4],
            need_bias=need[5])
        ctx.tape = None
        return d_feats, None, d_g if need[2] else None, d_b if need[3] else None, d_w, d_bias


class B200FeatureProjection(WavLMFeatureProjection):
    @classmethod
    def supports(cls, module) -> bool:
        return (isinstance(module, WavLMFeatureProjection) and tuple(module.projection.weight.shape) == (1024, 512)
                and module.projection.bias is not None)

    @classmethod
    def convert(cls, module: WavLMFeatureProjection) -> "B200FeatureProjection":
        if not cls.supports(module):
            raise ValueError("B200FeatureProjection implements LayerNorm(512) + Linear(512 -> 1024) with bias (wavlm-large)")
        module.__class__ = cls
        module._pk, module._pk_key = None, None
        return module

    def __init__(self, config):
        super().__init__(config)
        self._pk, self._pk_key = None, None
        if not self.supports(self):
            raise ValueError("B200FeatureProjection implements LayerNorm(512) + Linear(512 -> 1024) with bias (wavlm-large)")

    def _packs(self):
        """(bf16 weight, bf16 transposed weight), re-packed when the weight changed (see B200FeatureEncoder._weights_key)."""
        w = self.projection.weight
        key = (ops.param_generation(), w.data_ptr(), w._version)
        if getattr(self, "_pk_key", None) != key:
            with torch.no_grad():
                self._pk = ops.pack_feature_projection(w)
            self._pk_key = key
        return self._pk

    def forward(self, hidden_states: torch.Tensor):
        ln, pr = self.layer_norm, self.projection
        if torch.is_grad_enabled() and (hidden_states.requires_grad or any(
                p.requires_grad for p in (ln.weight, ln.bias, pr.weight, pr.bias))):
            hidden, norm = _FeatProjFn.apply(hidden_states, self, ln.weight, ln.bias, pr.weight, pr.bias)
        else:
            hidden, norm, _ = ops.feature_projection_fwd(hidden_states, ln.weight, ln.bias, ln.eps, self._packs()[0],
                                                         pr.bias, training=False)
        return self.dropout(hidden), norm
