"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/nrse_b200.h declares; host-only entry points behave (no compute call is made without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nrse_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nrse_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrse_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nrse_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(declared)


def test_version_and_errors(lib):
    assert lib.nrse_version() == 200
    assert lib.nrse_experiments_build() == 0   # the product library carries no results-corrupting timing hooks
    assert lib.nrse_strerror(0) == b"ok"
    assert b"invalid" in lib.nrse_strerror(-1)
    assert lib.nrse_mix_status_name(0) == b"ok"
    assert lib.nrse_mix_status_name(7) == b"scaled_noise_nan"
    with pytest.raises(_lib.NrseError):
        _lib.check(-2, "x")


def test_frontend_geometry_matches_hf_formula(lib):
    from nrse_b200.utils import synthetic
    for L in (400, 4000, 16000, 32000, 64000, 80000, 192000, 12345):
        T = (C.c_int32 * 7)()
        P = (C.c_int32 * 7)()
        assert lib.nrse_conv_frontend_geometry(L, T, P) == 0
        assert list(T) == synthetic.conv_out_lengths(L)
        for i in range(7):
            assert P[i] >= T[i]
            if i:
                assert P[i - 1] == 2 * P[i]
        assert P[6] - T[6] <= 1 + T[0] // 64 - T[6] + 1
    from nrse_b200 import ops
    for L in (400, 4000, 64000, 12345, 192000):
        assert ops._geometry_py(L) == ops.frontend_geometry(L)
    T = (C.c_int32 * 7)()
    P = (C.c_int32 * 7)()
    assert lib.nrse_conv_frontend_geometry(399, T, P) == -1  # empty output
    assert lib.nrse_conv_frontend_workspace_bytes(64, 64000) > 64 * 12800 * 512 * 2
    assert lib.nrse_conv_frontend_workspace_bytes(64, 100) == 0


def test_ema_chunk_planner(lib):
    ptr_t = np.array([0x1000, 0x90000, 0x200000], dtype=np.uint64)
    ptr_o = np.array([0x5000, 0xA0000, 0x300000], dtype=np.uint64)
    numel = np.array([8, 16384 * 2 + 4, 0], dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib.nrse_ema_plan_chunks_host(p(ptr_t), p(ptr_o), p(numel), 3, 16384, None, None, None, 0)
    assert n == 4
    ct = np.zeros(n, np.uint64); co = np.zeros(n, np.uint64); cn = np.zeros(n, np.int32)
    assert lib.nrse_ema_plan_chunks_host(p(ptr_t), p(ptr_o), p(numel), 3, 16384, p(ct), p(co), p(cn), n) == n
    assert cn.tolist() == [8, 16384, 16384, 4]
    assert ct.tolist() == [0x1000, 0x90000, 0x90000 + 65536, 0x90000 + 131072]
    assert co.tolist() == [0x5000, 0xA0000, 0xA0000 + 65536, 0xA0000 + 131072]
    # capacity too small / bad chunk size
    assert lib.nrse_ema_plan_chunks_host(p(ptr_t), p(ptr_o), p(numel), 3, 16384, p(ct), p(co), p(cn), 2) == -4
    assert lib.nrse_ema_plan_chunks_host(p(ptr_t), p(ptr_o), p(numel), 3, 1001, None, None, None, 0) == -1


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from nrse_b200 import ops
    with pytest.raises(_lib.NrseError):
        ops.mix_normalize(torch.zeros(2, 64), torch.zeros(2, 64), torch.zeros(2, dtype=torch.int32), [5.0])
    with pytest.raises(_lib.NrseError):
        ops.byol_loss(torch.zeros(2, 8), torch.zeros(2, 8))
    with pytest.raises(_lib.NrseError):
        ops.EmaPlan([torch.zeros(4)], [torch.zeros(4)])
