"""Drop-in replacement of HF ``WavLMPositionalConvEmbedding`` (hf:models/wavlm/modeling_wavlm.py:48-90): grouped
Conv1d(1024, 1024, k = 128, padding = 64, groups = 16) with weight-norm, last frame dropped, GELU -- the consumer of the
feature projection inside the WavLM encoder (``hidden_states + pos_conv_embed(hidden_states)``); SURVEY.md 8f-4.

It IS a ``WavLMPositionalConvEmbedding`` (same ``conv`` sub-module with its weight-norm parametrisation: parameter names
``conv.bias``, ``conv.parametrizations.weight.original0/1`` and shapes are untouched, checkpoints load as before); only
``forward`` differs: one cast launch + one tcgen05 implicit-GEMM launch on the channels-last hidden states (no transposes,
no padded copy: the TMA engine's out-of-bounds zero fill is the padding), and a native backward (GELU backward + bias
gradient, data-gradient GEMM with the tap-reversed weights, weight-gradient GEMM, weight-norm backward).

wavlm-large geometry only (hidden 1024, 16 groups, 128 taps, weight-norm through ``torch.nn.utils.parametrizations``);
other sizes keep the stock module (``supports`` says which).
"""
from __future__ import annotations

import torch
from transformers.models.wavlm.modeling_wavlm import WavLMPositionalConvEmbedding

from .. import ops


def _weight_norm_params(conv):
    par = getattr(conv, "parametrizations", None)
    if par is None or not hasattr(par, "weight"):
        return None
    return par.weight.original0, par.weight.original1   # g [1, 1, k], v [out, in / groups, k]


class _PosConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, g, v, bias):
        wf, wb, normsq = module._packs()
        y, xb, z = ops.pos_conv_fwd(x, wf, bias, training=True)
        ctx.save_for_backward(g, v)
        ctx.xb, ctx.z, ctx.wb, ctx.normsq = xb, z, wb, normsq
        return y

    @staticmethod
    def backward(ctx, g_y):
        if ctx.z is None:
            raise RuntimeError("B200PositionalConvEmbedding: the saved activations were freed by the first backward")
        g, v = ctx.saved_tensors
        need = ctx.needs_input_grad
        d_x, d_v, d_g, d_b = ops.pos_conv_bwd(g_y, ctx.xb, ctx.z, ctx.wb, v, g, ctx.normsq, need_x=need[0],
                                              need_w=need[2] or need[3], need_bias=need[4])
        ctx.xb = ctx.z = None
        return d_x, None, d_g if need[2] else None, d_v if need[3] else None, d_b


class B200PositionalConvEmbedding(WavLMPositionalConvEmbedding):
    @classmethod
    def supports(cls, module) -> bool:
        if not isinstance(module, WavLMPositionalConvEmbedding):
            return False
        conv = module.conv
        wn = _weight_norm_params(conv)
        return (wn is not None and conv.in_channels == 1024 and conv.out_channels == 1024 and conv.groups == 16
                and tuple(conv.kernel_size) == (128,) and tuple(conv.padding) == (64,) and conv.bias is not None
                and tuple(wn[0].shape) == (1, 1, 128) and getattr(module.padding, "num_pad_remove", 0) == 1)

    @classmethod
    def convert(cls, module: WavLMPositionalConvEmbedding) -> "B200PositionalConvEmbedding":
        if not cls.supports(module):
            raise ValueError("B200PositionalConvEmbedding implements Conv1d(1024, 1024, k=128, pad=64, groups=16) with "
                             "weight-norm over dim 2 (wavlm-large)")
        module.__class__ = cls
        module._pk, module._pk_key = None, None
        return module

    def _packs(self):
        g, v = _weight_norm_params(self.conv)
        key = (ops.param_generation(), g.data_ptr(), g._version, v.data_ptr(), v._version)
        if getattr(self, "_pk_key", None) != key:
            with torch.no_grad():
                self._pk = ops.pack_pos_conv(v, g)
            self._pk_key = key
        return self._pk

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        g, v = _weight_norm_params(self.conv)
        bias = self.conv.bias
        if torch.is_grad_enabled() and (hidden_states.requires_grad or g.requires_grad or v.requires_grad
                                        or bias.requires_grad):
            return _PosConvFn.apply(hidden_states, self, g, v, bias)
        return ops.pos_conv_fwd(hidden_states, self._packs()[0], bias, training=False)[0]
