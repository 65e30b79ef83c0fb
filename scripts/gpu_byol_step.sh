#!/bin/bash
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python scripts/train_step_bench.py --batch 64 --seconds 4 --steps 5 --warmup 2 --autocast > gpurun_out/${TAG:-r2}_byol_step_n1.log 2> gpurun_out/${TAG:-r2}_byol_step_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/train_step_bench.py --batch 64 --seconds 4 --steps 5 --warmup 2 --autocast > gpurun_out/${TAG:-r2}_byol_step_n$N.log 2> gpurun_out/${TAG:-r2}_byol_step_n$N.err
fi
tail -1 gpurun_out/${TAG:-r2}_byol_step_n$N.log | cut -c1-900
