"""Where does the exposed all-reduce time of the hot-path training step come from?  Times, at N ranks: the 1.3 GB fp32
all-reduce alone, the conv-frontend backward alone, and both started together (NCCL on its own stream).
torchrun --nproc-per-node N scripts/nccl_overlap_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from nrse_b200 import ops
from nrse_b200.utils import synthetic

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B, L = 64, 64000
layers = synthetic.frontend_weights("layer", seed=0)
w = [torch.from_numpy(l["conv"]).to(dev) for l in layers]
g = [torch.from_numpy(l["gamma"]).to(dev) for l in layers]
b = [torch.from_numpy(l["beta"]).to(dev) for l in layers]
x = torch.randn(B, L, device=dev)
T, P = ops.frontend_geometry(L)
gy = torch.randn(B, T[6], 512, device=dev)
packed = [ops.pack_conv_weight(t) for t in w[1:]]
dpacks = [ops.pack_conv_weight_dgrad(t) for t in w[1:]]
_, tape = ops.conv_frontend_train(x, w, g, b, "layer", packed=packed)
flat = torch.randn(325_434_048, device=dev) * 1e-4
bucket = 64 * 1024 * 1024
views = [flat[s:s + bucket] for s in range(0, flat.numel(), bucket)]


def bwd():
    ops.conv_frontend_backward(x, w, g, b, tape, gy, "layer", dgrad_packs=dpacks)


def ar(n=len(views)):
    works = [dist.all_reduce(v, op=dist.ReduceOp.AVG, async_op=True) for v in views[:n]]
    return works


def timed(fn, n=10):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def ar_alone():
    for wk in ar():
        wk.wait()


def both():
    works = ar()
    bwd()
    for wk in works:
        wk.wait()


res = {"world": world, "bytes": flat.numel() * 4}
res["allreduce_alone_ms"] = timed(ar_alone)
res["backward_alone_ms"] = timed(bwd)
res["both_ms"] = timed(both)
for reserve in (16, 32, 48):
    ops.set_sm_budget(148 - reserve)
    res[f"backward_alone_reserve{reserve}_ms"] = timed(bwd)
    res[f"both_reserve{reserve}_ms"] = timed(both)
ops.set_sm_budget(148)
res["busbw_gbs_alone"] = 2 * (world - 1) / world * res["bytes"] / (res["allreduce_alone_ms"] * 1e-3) / 1e9

# the same with this repository's NVSwitch-multicast kernel (GradArena, csrc/allreduce.cu)
from nrse_b200.train import GradArena
prm = [torch.nn.Parameter(torch.empty(0, device=dev))]
big = torch.nn.Parameter(torch.zeros(flat.numel(), device=dev))
for ctas in (16, 32, 64, 148):
    arena = GradArena([[big]], multimem=True, multimem_ctas=ctas)
    arena.flat.normal_(0, 1e-4)

    def mm_alone():
        arena.all_reduce_async(0)
        arena.wait()

    def mm_both():
        arena.all_reduce_async(0)
        bwd()
        arena.wait()

    res[f"multimem_alone_ctas{ctas}_ms"] = timed(mm_alone)
    res[f"multimem_both_ctas{ctas}_ms"] = timed(mm_both)
    del arena
    big.grad = None
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
