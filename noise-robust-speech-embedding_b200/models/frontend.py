"""Drop-in replacement of HF ``WavLMFeatureEncoder`` (hf:models/wavlm/modeling_wavlm.py:754-789) whose forward runs
the B200 kernels.  It IS a ``WavLMFeatureEncoder`` (same ``conv_layers`` ModuleList, same parameter names/shapes,
``_freeze_parameters``, ``_requires_grad``, ``gradient_checkpointing``), so checkpoints, ``named_parameters()``
substring matches (ref:src/models/emotion.py:126-129) and ``load_state_dict`` keep working; only ``forward`` differs.

forward([B,L] fp32) -> [B,512,T] fp32, returned as a transposed VIEW of the kernels' channels-last [B,T,512]
output -- ``WavLMModel.forward`` transposes it straight back (hf:...:1061), so no copy is ever made.

Backward: in LayerNorm mode (wavlm-large) the training forward keeps a tape (bf16 activations, normalised
pre-affine values, 1/std per frame) and the backward runs the native kernels -- LayerNorm+GELU backward, tcgen05
weight-gradient GEMM with split-K, data-gradient GEMMs (``ops.conv_frontend_backward``).  GroupNorm mode (wavlm-base)
and the rarely needed waveform gradient fall back to recomputation with stock torch ops (``_FrontendFn``).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F
from transformers.models.wavlm.modeling_wavlm import WavLMFeatureEncoder

from .. import ops


def _torch_stack(x: torch.Tensor, conv_w, gammas, betas, norm_mode: str) -> torch.Tensor:
    """Stock-torch restatement of the conv stack, used ONLY to differentiate (recompute in backward)."""
    h = x[:, None]
    for i, w in enumerate(conv_w):
        h = F.conv1d(h, w, stride=ops.CONV_STRIDE[i])
        if norm_mode == "layer":
            h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), gammas[i], betas[i], 1e-5).transpose(1, 2)
        elif i == 0:
            h = F.group_norm(h, h.shape[1], gammas[0], betas[0], 1e-5)
        h = F.gelu(h)
    return h


class _FrontendNativeFn(torch.autograd.Function):
    """LayerNorm-mode frontend with the native backward kernels (training forward keeps a tape of activations)."""

    @staticmethod
    def forward(ctx, x, packed_holder, *params):
        conv_w, gammas, betas = params[:7], params[7:14], params[14:21]
        y, tape = ops.conv_frontend_train(x, conv_w, list(gammas), list(betas), packed=packed_holder())
        ctx.save_for_backward(x, *params)
        ctx.tape = tape
        return y.transpose(1, 2)

    @staticmethod
    def backward(ctx, grad_out):
        x, *params = ctx.saved_tensors
        conv_w, gammas, betas = params[:7], params[7:14], params[14:21]
        dw, dg, db = ops.conv_frontend_backward(x, conv_w, list(gammas), list(betas), ctx.tape, grad_out.transpose(1, 2))
        ctx.tape = None
        grads = [*dw, *dg, *db]
        need = ctx.needs_input_grad[2:]
        # the waveform gradient is not produced (nothing upstream of the waveform is trainable in the reference)
        return (None, None, *[g if n else None for g, n in zip(grads, need)])


class _FrontendFn(torch.autograd.Function):
    """Fallback used for GroupNorm mode (wavlm-base): gradients by recomputation with stock torch ops."""

    @staticmethod
    def forward(ctx, x, norm_mode, n_norm, packed_holder, *params):
        conv_w = params[:7]
        gammas, betas = params[7:7 + n_norm], params[7 + n_norm:7 + 2 * n_norm]
        y = ops.conv_frontend(x, conv_w, list(gammas), list(betas), norm_mode, out_dtype=torch.float32,
                              packed=packed_holder())
        ctx.save_for_backward(x, *params)
        ctx.norm_mode, ctx.n_norm = norm_mode, n_norm
        return y.transpose(1, 2)

    @staticmethod
    def backward(ctx, grad_out):
        x, *params = ctx.saved_tensors
        n_norm = ctx.n_norm
        need = [i for i, g in enumerate(ctx.needs_input_grad[4:]) if g]
        with torch.enable_grad():
            ps = [p.detach().requires_grad_(i in need) for i, p in enumerate(params)]
            xin = x.detach().requires_grad_(ctx.needs_input_grad[0])
            y = _torch_stack(xin, ps[:7], ps[7:7 + n_norm], ps[7 + n_norm:], ctx.norm_mode)
            wrt = ([xin] if ctx.needs_input_grad[0] else []) + [ps[i] for i in need]
            grads = torch.autograd.grad(y, wrt, grad_out, allow_unused=True)
        grads = list(grads)
        gx = grads.pop(0) if ctx.needs_input_grad[0] else None
        out: List[Optional[torch.Tensor]] = [None] * len(params)
        for i, g in zip(need, grads):
            out[i] = g
        return (gx, None, None, None, *out)


class B200FeatureEncoder(WavLMFeatureEncoder):
    """``WavLMFeatureEncoder`` with the sm_100a forward.  Build one with ``B200FeatureEncoder(config)`` or convert an
    existing HF module in place with ``B200FeatureEncoder.convert(module)`` (keeps its parameters)."""

    out_dtype = torch.float32

    @classmethod
    def convert(cls, module: WavLMFeatureEncoder) -> "B200FeatureEncoder":
        if not isinstance(module, WavLMFeatureEncoder):
            raise TypeError(f"expected a WavLMFeatureEncoder, got {type(module).__name__}")
        kernels = [tuple(l.conv.kernel_size)[0] for l in module.conv_layers]
        strides = [tuple(l.conv.stride)[0] for l in module.conv_layers]
        if tuple(kernels) != ops.CONV_KERNEL or tuple(strides) != ops.CONV_STRIDE:
            raise ValueError("the B200 frontend implements the WavLM/wav2vec2 geometry k=(10,3,3,3,3,2,2), s=(5,2,...)")
        if any(l.conv.bias is not None for l in module.conv_layers):
            raise ValueError("conv_bias=True is not supported (WavLM uses bias-free convolutions)")
        module.__class__ = cls
        module._packed, module._packed_key = None, None
        return module

    def __init__(self, config):
        super().__init__(config)
        self._packed, self._packed_key = None, None

    # -- helpers ---------------------------------------------------------------------------------------------------
    @property
    def norm_mode(self) -> str:
        return "layer" if hasattr(self.conv_layers[1], "layer_norm") else "group"

    def _params(self):
        conv_w = [l.conv.weight for l in self.conv_layers]
        n_norm = 7 if self.norm_mode == "layer" else 1
        gammas = [self.conv_layers[i].layer_norm.weight for i in range(n_norm)]
        betas = [self.conv_layers[i].layer_norm.bias for i in range(n_norm)]
        return conv_w, gammas, betas, n_norm

    def _packed_weights(self):
        """bf16 [512, k*512] copies of conv weights 1..6, re-packed only when a weight changed (version counter)."""
        ws = [l.conv.weight for l in self.conv_layers[1:]]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if getattr(self, "_packed_key", None) != key:
            with torch.no_grad():
                self._packed = [ops.pack_conv_weight(w) for w in ws]
            self._packed_key = key
        return self._packed

    def forward(self, input_values: torch.Tensor) -> torch.Tensor:
        conv_w, gammas, betas, n_norm = self._params()
        x = input_values
        if x.dim() == 3:
            x = x.squeeze(1)
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(
            p.requires_grad for p in (*conv_w, *gammas, *betas)))
        if needs_grad:
            if self.norm_mode == "layer" and not x.requires_grad:
                return _FrontendNativeFn.apply(x.float(), self._packed_weights, *conv_w, *gammas, *betas)
            return _FrontendFn.apply(x.float(), self.norm_mode, n_norm, self._packed_weights, *conv_w, *gammas, *betas)
        y = ops.conv_frontend(x, conv_w, gammas, betas, self.norm_mode, out_dtype=self.out_dtype,
                              packed=self._packed_weights())
        return y.transpose(1, 2)
