#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  scripts/nccl_overlap_probe.py > gpurun_out/${TAG:-r2}_probe_n$N.log 2>&1
grep -i "nvls\|channels\|Algo\|Proto" gpurun_out/${TAG:-r2}_probe_n$N.log | head -12
tail -1 gpurun_out/${TAG:-r2}_probe_n$N.log
