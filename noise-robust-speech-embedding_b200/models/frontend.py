"""Drop-in replacement of HF ``WavLMFeatureEncoder`` (hf:models/wavlm/modeling_wavlm.py:754-789) whose forward AND
backward run the B200 kernels.  It IS a ``WavLMFeatureEncoder`` (same ``conv_layers`` ModuleList, same parameter
names/shapes, ``_freeze_parameters``, ``_requires_grad``, ``gradient_checkpointing``), so checkpoints,
``named_parameters()`` substring matches (ref:src/models/emotion.py:126-129) and ``load_state_dict`` keep working; only
``forward`` differs.

forward([B,L] fp32) -> [B,512,T] fp32, returned as a transposed VIEW of the kernels' channels-last [B,T,512]
output -- ``WavLMModel.forward`` transposes it straight back (hf:...:1061), so no copy is ever made.

Backward (both norm modes: LayerNorm on every layer = wavlm-large, GroupNorm on layer 0 only = wavlm-base): the training
forward keeps a tape (bf16 activations, normalised pre-affine values, statistics) and the backward runs the native
kernels (``ops.conv_frontend_backward``): norm + GELU backward, tcgen05 weight-gradient GEMMs with split-K, data-gradient
GEMMs.  Only the gradients autograd actually asks for are computed -- a partially unfrozen encoder
(ref:src/models/emotion.py:114-129) stops at its lowest trainable conv layer.  The waveform gradient is not implemented
(nothing upstream of the waveform is trainable anywhere in the reference) and asking for it raises.

The bf16 weight packs the kernels consume are cached and keyed on (parameter address, ``tensor._version``,
``ops.param_generation()``): torch-side writers bump ``_version``; the multi-tensor kernels (``FusedAdamWEma``,
``EmaPlan``) write through raw pointers and bump the generation counter instead.
"""
from __future__ import annotations

import torch
from transformers.models.wavlm.modeling_wavlm import WavLMFeatureEncoder

from .. import ops


class _FrontendFn(torch.autograd.Function):
    """Training forward (keeps a tape) + native backward, both norm modes."""

    @staticmethod
    def forward(ctx, x, module, norm_mode, n_norm, *params):
        conv_w, gammas, betas = params[:7], params[7:7 + n_norm], params[7 + n_norm:7 + 2 * n_norm]
        y, tape = ops.conv_frontend_train(x, conv_w, list(gammas), list(betas), norm_mode, packed=module._packed_weights())
        ctx.save_for_backward(x, *params)
        ctx.tape, ctx.module, ctx.norm_mode, ctx.n_norm = tape, module, norm_mode, n_norm
        return y.transpose(1, 2)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.tape is None:
            raise RuntimeError("B200FeatureEncoder: the tape of this forward was freed by its first backward "
                               "(retain_graph / a second backward through the conv frontend is not supported)")
        x, *params = ctx.saved_tensors
        n_norm = ctx.n_norm
        conv_w, gammas, betas = params[:7], params[7:7 + n_norm], params[7 + n_norm:7 + 2 * n_norm]
        need = ctx.needs_input_grad[4:]
        need_w, need_g, need_b = need[:7], need[7:7 + n_norm], need[7 + n_norm:7 + 2 * n_norm]
        need_aff = [g or b for g, b in zip(need_g, need_b)]
        lowest = min([i for i, v in enumerate(need_w) if v] + [i for i, v in enumerate(need_aff) if v], default=7)
        dw, dg, db = ops.conv_frontend_backward(x, conv_w, list(gammas), list(betas), ctx.tape, grad_out.transpose(1, 2),
                                                ctx.norm_mode, dgrad_packs=ctx.module._dgrad_packs(lowest),
                                                need_w=need_w, need_affine=need_aff)
        ctx.tape = None
        grads = [*dw, *[g if n else None for g, n in zip(dg, need_g)], *[b if n else None for b, n in zip(db, need_b)]]
        return (None, None, None, None, *grads)


class B200FeatureEncoder(WavLMFeatureEncoder):
    """``WavLMFeatureEncoder`` with the sm_100a forward and backward.  Build one with ``B200FeatureEncoder(config)`` or
    convert an existing HF module in place with ``B200FeatureEncoder.convert(module)`` (keeps its parameters)."""

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
        module._reset_packs()
        return module

    def __init__(self, config):
        super().__init__(config)
        self._reset_packs()

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _reset_packs(self) -> None:
        self._packed, self._packed_key = None, None
        self._dpacks, self._dpacks_key = None, None

    @property
    def norm_mode(self) -> str:
        return "layer" if hasattr(self.conv_layers[1], "layer_norm") else "group"

    def _params(self):
        conv_w = [l.conv.weight for l in self.conv_layers]
        n_norm = 7 if self.norm_mode == "layer" else 1
        gammas = [self.conv_layers[i].layer_norm.weight for i in range(n_norm)]
        betas = [self.conv_layers[i].layer_norm.bias for i in range(n_norm)]
        return conv_w, gammas, betas, n_norm

    def _weights_key(self):
        """Changes whenever a conv weight of layers 1..6 may have changed: torch-side in-place writes bump ``_version``;
        the raw-pointer writers (fused optimizer / EMA kernels) bump ``ops.param_generation()``."""
        ws = [l.conv.weight for l in self.conv_layers[1:]]
        return (ops.param_generation(), tuple((w.data_ptr(), w._version) for w in ws))

    def _packed_weights(self):
        """bf16 [512, k*512] copies of conv weights 1..6 (forward operands), re-packed when a weight changed."""
        key = self._weights_key()
        if getattr(self, "_packed_key", None) != key:
            with torch.no_grad():
                self._packed = [ops.pack_conv_weight(l.conv.weight) for l in self.conv_layers[1:]]
            self._packed_key = key
        return self._packed

    def _dgrad_packs(self, lowest: int = 0):
        """Data-gradient operands of layers 1..6 (same cache key); layers at or below ``lowest`` never propagate a data
        gradient and get ``None``."""
        key = (self._weights_key(), lowest)
        if getattr(self, "_dpacks_key", None) != key:
            with torch.no_grad():
                self._dpacks = [ops.pack_conv_weight_dgrad(l.conv.weight) if i > lowest else None
                                for i, l in enumerate(self.conv_layers) if i >= 1]
            self._dpacks_key = key
        return self._dpacks

    def forward(self, input_values: torch.Tensor) -> torch.Tensor:
        conv_w, gammas, betas, n_norm = self._params()
        x = input_values
        if x.dim() == 3:
            x = x.squeeze(1)
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("B200FeatureEncoder does not compute the waveform gradient (the reference never "
                                      "trains anything upstream of the waveform); detach the input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in (*conv_w, *gammas, *betas)):
            return _FrontendFn.apply(x.float(), self, self.norm_mode, n_norm, *conv_w, *gammas, *betas)
        y = ops.conv_frontend(x, conv_w, gammas, betas, self.norm_mode, out_dtype=self.out_dtype,
                              packed=self._packed_weights())
        return y.transpose(1, 2)
