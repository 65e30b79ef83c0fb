"""N-rank check of GradArena's NVSwitch-multicast all-reduce (csrc/allreduce.cu) against the exact mean and against NCCL.
torchrun --nproc-per-node N scripts/multimem_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from nrse_b200.train import GradArena

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
sizes_a = [7, 1000, 33 * 7, 1 << 20, 5, (1 << 22) + 12]
sizes_b = [512 * 512 * 3, 512, 10]
res = {"world": world}
for impl in (True, False):
    pa = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in sizes_a]
    pb = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in sizes_b]
    arena = GradArena([pa, pb], bucket_bytes=1 << 20, multimem=impl)
    g = torch.Generator(device=dev).manual_seed(1234)
    base = [torch.randn(p.numel(), device=dev, generator=g) for p in pa + pb]     # same on every rank
    want = []
    for p, b0 in zip(pa + pb, base):
        p.grad.copy_(b0 * (rank + 1))                                              # rank-dependent gradients
        want.append(b0.double() * (sum(range(1, world + 1)) / world))             # their exact mean
    guard = arena.flat.clone()
    arena.all_reduce_async(0)
    torch.cuda.current_stream().synchronize()  # group 1 must still be untouched while group 0 travels
    arena.all_reduce_async(1)
    arena.wait()
    torch.cuda.synchronize()
    err = max(float((p.grad.double() - w).abs().max() / w.abs().max()) for p, w in zip(pa + pb, want))
    # the 4-element alignment padding between tensors must stay as it was (zeros)
    res["multimem" if impl else "nccl"] = {"impl": arena.impl, "max_rel_err": err}
    assert err < 1e-6, (impl, err)
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
