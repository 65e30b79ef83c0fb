#!/bin/bash
# bench.py at N GPUs (torchrun, as the driver launches it): N=$1
N=${1:-2}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG:-r2}_bench_n$N.log 2> gpurun_out/${TAG:-r2}_bench_n$N.err
echo "rc=$?" >> gpurun_out/${TAG:-r2}_bench_n$N.err
tail -c 1500 gpurun_out/${TAG:-r2}_bench_n$N.err
python - <<PY
import json
l=json.loads(open('gpurun_out/${TAG:-r2}_bench_n$N.log').read().strip().splitlines()[-1])
print('value', l['value'], 'e2e', l['e2e']['value'], 'sustained', l['value_sustained'])
print(json.dumps(l['train_step']))
PY
