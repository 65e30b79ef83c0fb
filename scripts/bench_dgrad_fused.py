"""A/B of the data gradient + LayerNorm / GELU backward of one layer pair at the BASELINE shape: two kernels against
the fused epilogue (nrse_conv_layer_dgrad_lnbwd).  --layer i = the layer whose data gradient runs (1..6)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops

dev = torch.device("cuda:0")
B, L = 64, 64000
T, P = ops.frontend_geometry(L)
KS = (10, 3, 3, 3, 3, 2, 2)
layers = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 3]
once = "--once" in sys.argv   # a single launch of each (for ncu)

def timeit(fn, n=10):
    if once:
        fn(); torch.cuda.synchronize(); return 0.0
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3

for i in layers:
    k = KS[i]
    rows_out, rows_prev = B * P[i], B * P[i - 1]
    g = torch.Generator(device=dev).manual_seed(i)
    dz = (torch.randn(rows_out, 512, device=dev, generator=g)).bfloat16()
    xhat = torch.randn(rows_prev, 512, device=dev, generator=g).bfloat16()
    rstd = torch.rand(rows_prev, device=dev, generator=g) + 0.5
    gamma = 1 + 0.1 * torch.randn(512, device=dev, generator=g)
    beta = 0.1 * torch.randn(512, device=dev, generator=g)
    w = torch.randn(512, 512, k, device=dev, generator=g) * (2.0 / (512 * k)) ** 0.5
    even, odd = ops.pack_conv_weight_dgrad(w)
    lib = ops._lib.load()
    dx = torch.empty(rows_prev, 512, dtype=torch.bfloat16, device=dev)
    dg = torch.zeros(512, device=dev); db = torch.zeros(512, device=dev)
    st = ops._stream
    def plain():
        ops.check(lib.nrse_conv_layer_dgrad(ops._ptr(dz), rows_out, ops._ptr(even), ops._ptr(odd), k, ops._ptr(dx), st()), "d")
    def ln():
        ops.check(lib.nrse_ln_gelu_bwd(ops._ptr(dx), ops._dtype_code(dx), P[i - 1], ops._ptr(xhat), ops._ptr(rstd), ops._ptr(gamma),
                                       ops._ptr(beta), ops._ptr(dx), ops._ptr(dg), ops._ptr(db), rows_prev, P[i - 1], T[i - 1], st()), "l")
    def fused(aff=True):
        ops.check(lib.nrse_conv_layer_dgrad_lnbwd(ops._ptr(dz), rows_out, ops._ptr(even), ops._ptr(odd), k, ops._ptr(xhat),
                                                  ops._ptr(rstd), ops._ptr(gamma), ops._ptr(beta), ops._ptr(dx),
                                                  ops._ptr(dg) if aff else None, ops._ptr(db) if aff else None,
                                                  P[i - 1], T[i - 1], st()), "f")
    res = {"layer": i, "dgrad_us": timeit(plain), "ln_gelu_bwd_us": timeit(ln), "fused_us": timeit(fused),
           "fused_no_affine_us": timeit(lambda: fused(False))}
    res["two_kernels_us"] = res["dgrad_us"] + res["ln_gelu_bwd_us"]
    print(json.dumps(res), flush=True)
