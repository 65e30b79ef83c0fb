"""Device time of the conv feature encoder training forward + native backward at the BASELINE shape (64 x 4 s)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrse_b200 import ops
from nrse_b200.utils import synthetic

dev = torch.device("cuda:0")
B, L = 64, 64000
layers = synthetic.frontend_weights("layer", seed=0)
x = synthetic.waveforms(B, L, seed=1)[0]
x = torch.from_numpy(((x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)).astype("float32")).to(dev)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
packed = [ops.pack_conv_weight(t) for t in w[1:]]
dpacks = [ops.pack_conv_weight_dgrad(t) for t in w[1:]]
T, P = ops.frontend_geometry(L)
gy = torch.randn(B, T[6], 512, device=dev)

def timeit(fn, n=5, repeats=3):
    """best of `repeats` averages over n eager calls (the calls allocate their outputs -- 3.3 GB of tape for the training
    forward --, so single averages wobble by tens of per cent with the allocator / power state)"""
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best, out

t_fwd, _ = timeit(lambda: ops.conv_frontend(x, w, g, b, "layer", out_dtype=torch.bfloat16, packed=packed))
t_train, (y, tape) = timeit(lambda: ops.conv_frontend_train(x, w, g, b, "layer", packed=packed))
bwd = lambda: ops.conv_frontend_backward(x, w, g, b, tape, gy, "layer", dgrad_packs=dpacks)
if "--timeline" in sys.argv:   # CUPTI timeline of one warm backward of each form: start, duration, gap to the previous kernel
    from torch.profiler import profile, ProfilerActivity
    for on in (False, True):
        ops.set_bwd_fusion(on)
        for _ in range(3):
            bwd()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            bwd(); torch.cuda.synchronize()
        ev = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
        t0 = ev[0].time_range.start
        print("fusion", on, "span_us", ev[-1].time_range.end - t0, "sum_us", sum(e.time_range.end - e.time_range.start for e in ev))
        prev_end = t0
        for e in ev:
            print(f"  {e.time_range.start - t0:8.1f} dur {e.time_range.end - e.time_range.start:7.1f} gap {e.time_range.start - prev_end:6.1f}  {e.name[:60]}")
            prev_end = max(prev_end, e.time_range.end)
    sys.exit(0)
if "--profile" in sys.argv:   # under ncu: two backward passes of each form (the second of each is the warm one)
    for on in (False, False, True, True):
        ops.set_bwd_fusion(on); bwd(); torch.cuda.synchronize()
    sys.exit(0)
ops.set_bwd_fusion(False)
t_bwd_sep, _ = timeit(bwd)
ops.set_bwd_fusion(True)      # LayerNorm / GELU backward inside the data-gradient epilogue
t_bwd_fused, _ = timeit(bwd)
ops.set_bwd_fusion(False)
t_bwd_sep = min(t_bwd_sep, timeit(bwd)[0])
ops.set_bwd_fusion(ops.DEFAULT_BWD_FUSION)
t_bwd = t_bwd_fused if ops.DEFAULT_BWD_FUSION else t_bwd_sep
if "--quick" in sys.argv:
    print(json.dumps({"shape": [B, L], "fwd_ms": t_fwd, "train_fwd_ms": t_train, "bwd_ms": t_bwd,
                      "bwd_separate_ms": t_bwd_sep, "bwd_fused_ms": t_bwd_fused}))
    sys.exit(0)
fwd_flops = sum(2.0 * B * T[i] * 512 * 512 * k for i, k in enumerate((10, 3, 3, 3, 3, 2, 2)) if i > 0)
# stock torch (cuDNN / ATen) forward+backward of the same stack for comparison, fp32 and bf16 autocast
import torch.nn.functional as F
def torch_stack(xx, ws, gs, bs):
    h = xx[:, None]
    for i, wt in enumerate(ws):
        h = F.conv1d(h, wt, stride=ops.CONV_STRIDE[i])
        h = F.layer_norm(h.transpose(1, 2), (512,), gs[i], bs[i], 1e-5).transpose(1, 2)
        h = F.gelu(h)
    return h
ws = [t.clone().requires_grad_(True) for t in w]; gs = [t.clone().requires_grad_(True) for t in g]; bs = [t.clone().requires_grad_(True) for t in b]
res = {}
for name, ac in (("fp32", False), ("bf16_autocast", True)):
    def fb():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            out = torch_stack(x, ws, gs, bs)
        out.backward(gy.transpose(1, 2).to(out.dtype))
    def fo():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            return torch_stack(x, ws, gs, bs)
    res[name] = {"fwd_ms": timeit(fo, 3, 2)[0], "fwd_bwd_ms": timeit(fb, 3, 2)[0]}
print(json.dumps({"shape": [B, L], "fwd_ms": t_fwd, "train_fwd_ms": t_train, "bwd_ms": t_bwd,

                  "bwd_tflops_2x_fwd_gemm": 2 * fwd_flops / (t_bwd * 1e-3) / 1e12,
                  "stock_torch_on_this_gpu": res}))
