#!/bin/bash
# N-GPU pass: bench.py as the driver launches it, then BASELINE configs[2] (full BYOL step), [3] (emotion fine-tune step), [4] (eval sweep)
N=${1:-8}
TAG=${TAG:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.log 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
timeout 600 $TR --master-port 29512 scripts/train_step_bench.py --batch 64 --seconds 4 --steps 5 --warmup 2 --autocast > gpurun_out/${TAG}_byol_step_n$N.log 2> gpurun_out/${TAG}_byol_step_n$N.err; echo "byol rc=$?"
timeout 600 $TR --master-port 29513 scripts/emotion_step_bench.py > gpurun_out/${TAG}_emotion_step_n$N.log 2> gpurun_out/${TAG}_emotion_step_n$N.err; echo "emotion rc=$?"
timeout 600 $TR --master-port 29514 scripts/eval_sweep_bench.py > gpurun_out/${TAG}_eval_sweep_n$N.log 2> gpurun_out/${TAG}_eval_sweep_n$N.err; echo "eval rc=$?"
for f in bench byol_step emotion_step eval_sweep; do echo "== $f"; tail -n 3 gpurun_out/${TAG}_${f}_n$N.log | cut -c1-700; tail -n 2 gpurun_out/${TAG}_${f}_n$N.err | cut -c1-300; done
