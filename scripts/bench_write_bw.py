"""Pure-write, pure-read and copy bandwidth of the B200's HBM with stock torch kernels (fill_, sum, copy_) over 2 GiB:
the ceiling an output-dominated kernel (layer 0 writes 839 MB and reads 16 MB) can be held against."""
import torch
dev = torch.device("cuda:0")
n = 1 << 29  # fp32 elements = 2 GiB
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
tw = t(lambda: a.fill_(1.0))
tr = t(lambda: a.sum())
tc = t(lambda: b.copy_(a))
x16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
tcv = t(lambda: x16.copy_(a))   # read 4 B, write 2 B per element
print(f"fill_ (write only): {4*n/tw/1e9:.0f} GB/s   sum (read only): {4*n/tr/1e9:.0f} GB/s   copy_ (read+write): {8*n/tc/1e9:.0f} GB/s   "
      f"fp32->bf16 convert (4 B read + 2 B write): {6*n/tcv/1e9:.0f} GB/s", flush=True)
