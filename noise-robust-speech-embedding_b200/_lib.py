"""ctypes binding of libnrse_b200.so, the C-ABI library declared in include/nrse_b200.h.

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
``load()`` never compiles anything; ``__graft_entry__.build()`` (or ``python csrc/build.py``) does.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NRSE_B200_LIB: another build of the SAME library (A/B timing of kernel changes on one box, scripts/ only)
LIB_PATH = os.environ.get("NRSE_B200_LIB") or os.path.join(_HERE, "csrc", "libnrse_b200.so")

NRSE_OK = 0
DTYPE_F32 = 0
DTYPE_BF16 = 1
NORM_LAYER = 0
NORM_GROUP = 1
N_LAYERS = 7
CHANNELS = 512


class NrseError(RuntimeError):
    pass


class FrontendParams(C.Structure):
    """struct nrse_frontend_params (include/nrse_b200.h)."""
    _fields_ = [
        ("w0", C.c_void_p),
        ("w_packed", C.c_void_p * (N_LAYERS - 1)),
        ("gamma", C.c_void_p * N_LAYERS),
        ("beta", C.c_void_p * N_LAYERS),
    ]


class FrontendBwdWeights(C.Structure):
    """struct nrse_frontend_bwd_weights."""
    _fields_ = [("wt_even", C.c_void_p * (N_LAYERS - 1)), ("wt_odd", C.c_void_p * (N_LAYERS - 1))]


class TensorCheck(C.Structure):
    """struct nrse_tensor_check (64 bytes)."""
    _fields_ = [("flags", C.c_int32), ("abs_max", C.c_float), ("max", C.c_float), ("min", C.c_float),
                ("abs_sum", C.c_double), ("sum", C.c_double), ("sumsq", C.c_double), ("numel", C.c_int64),
                ("reserved", C.c_int64), ("reserved2", C.c_int64)]


class FrontendGrads(C.Structure):
    """struct nrse_frontend_grads."""
    _fields_ = [("dw0", C.c_void_p), ("dw", C.c_void_p * (N_LAYERS - 1)), ("dgamma", C.c_void_p * N_LAYERS),
                ("dbeta", C.c_void_p * N_LAYERS)]


_i, _i64, _p, _sz, _f = C.c_int, C.c_int64, C.c_void_p, C.c_size_t, C.c_float

# name -> (restype, argtypes); every symbol of include/nrse_b200.h
SIGNATURES = {
    "nrse_version": (_i, []),
    "nrse_strerror": (C.c_char_p, [_i]),
    "nrse_last_cuda_error": (_i, []),
    "nrse_check_device": (_i, []),
    "nrse_experiments_build": (_i, []),
    "nrse_check_tensors_f32": (_i, [_p, _p, _i, _f, _f, _p, _p]),
    "nrse_mix_normalize_f32": (_i, [_p, _p, _p, C.POINTER(C.c_double), _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "nrse_mix_normalize_retry_f32": (_i, [_p, _p, _p, C.POINTER(C.c_double), _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "nrse_mix_substitute_rows_f32": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "nrse_mix_batch_f32": (_i, [_p, _p, _p, C.POINTER(C.c_double), _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "nrse_mix_status_name": (C.c_char_p, [_i]),
    "nrse_mix_set_variant": (_i, [_i]),
    "nrse_mix_set_cluster": (_i, [_i]),
    "nrse_mix_set_carveout": (_i, [_i]),
    "nrse_ema_plan_chunks_host": (_i64, [_p, _p, _p, _i, _i64, _p, _p, _p, _i64]),
    "nrse_ema_chunks_f32": (_i, [_p, _p, _p, _i64, _f, _f, _p]),
    "nrse_optim_plan_chunks_host": (_i64, [_p, _p, _p, _p, _p, _p, _i, _i64, _p, _p, _i64]),
    "nrse_optim_partials_count": (_i, []),
    "nrse_grad_sqnorm_chunks_f32": (_i, [_p, _p, _i64, _p, _p]),
    "nrse_clip_adamw_ema_chunks_f32": (_i, [_p, _i64, _p, _i64, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_double, _i64, C.c_double, C.c_double, _p, _i, _p, _p]),
    "nrse_asp_pool_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "nrse_asp_pool_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "nrse_asp_pool_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _i, _i, _i, _p]),
    "nrse_byol_loss_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "nrse_byol_loss_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "nrse_conv_frontend_geometry": (_i, [_i, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "nrse_conv_frontend_workspace_bytes": (_sz, [_i, _i]),
    "nrse_conv_frontend_pack_weights": (_i, [_p, _p, _i, _p]),
    "nrse_conv_frontend_fwd": (_i, [_p, C.POINTER(FrontendParams), _i, _p, _i, _p, _sz, _p, _i, _i, _p]),
    "nrse_conv_layer0_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "nrse_conv_layer_fwd": (_i, [_p, _i64, _p, _i, _i, _p, _p, _p, _i, _i64, _p]),
    "nrse_conv_frontend_set_variant": (_i, [_i]),
    "nrse_conv_frontend_set_layer0_variant": (_i, [_i]),
    "nrse_conv_frontend_set_tile_order": (_i, [_i]),
    "nrse_conv_frontend_set_bwd_fusion": (_i, [_i]),
    "nrse_conv_frontend_set_sm_budget": (_i, [_i]),
    "nrse_conv_frontend_set_l2_prefetch": (_i, [_i]),
    "nrse_conv_frontend_tape_bytes": (_sz, [_i, _i]),
    "nrse_conv_frontend_fwd_train": (_i, [_p, C.POINTER(FrontendParams), _i, _p, _i, _p, _sz, _i, _i, _p]),
    "nrse_conv_frontend_pack_weights_dgrad": (_i, [_p, _p, _p, _i, _p]),
    "nrse_conv_frontend_bwd_workspace_bytes": (_sz, [_i, _i]),
    "nrse_conv_frontend_bwd": (_i, [_p, C.POINTER(FrontendParams), C.POINTER(FrontendBwdWeights), _i, _p, _p, _i,
                                    C.POINTER(FrontendGrads), _p, _sz, _i, _i, _p]),
    "nrse_ln_gelu_bwd": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _p]),
    "nrse_conv_layer0_wgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "nrse_conv_layer0_gn_bwd_scratch_bytes": (_sz, [_i]),
    "nrse_conv_layer0_gn_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "nrse_conv_layer_wgrad": (_i, [_p, _p, _i64, _i, _p, _i, _p]),
    "nrse_conv_layer_dgrad": (_i, [_p, _i64, _p, _p, _i, _p, _p]),
    "nrse_conv_layer_dgrad_lnbwd": (_i, [_p, _i64, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "nrse_multimem_allreduce_mean_f32": (_i, [_p, _i64, _i64, _i, _i, _i, _p]),
    "nrse_pos_conv_pack_bytes": (_sz, []),
    "nrse_pos_conv_pack": (_i, [_p, _p, _p, _p, _p, _p]),
    "nrse_pos_conv_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "nrse_pos_conv_bwd_workspace_bytes": (_sz, [_i, _i]),
    "nrse_pos_conv_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "nrse_feature_projection_pack": (_i, [_p, _p, _p, _p]),
    "nrse_feature_projection_tape_bytes": (_sz, [_i64]),
    "nrse_feature_projection_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p, _p, _p, _i, _p]),
    "nrse_feature_projection_bwd_workspace_bytes": (_sz, [_i64]),
    "nrse_feature_projection_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NrseError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                "(nrse_b200 has no CPU or pure-PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != NRSE_OK:
        lib = load()
        msg = lib.nrse_strerror(rc).decode()
        extra = ""
        if rc == -3:
            extra = f" (cudaError {lib.nrse_last_cuda_error()})"
        raise NrseError(f"{what or 'nrse call'} failed: {msg}{extra}")
