#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-2} gpurun_out/$name.log | cut -c1-1200; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run bench_n2 900 $TR bench.py --gpus 2 --steps 100 --warmup 3
run bench_ref_n2 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1
run byol_n2 900 $TR scripts/train_step_bench.py --autocast --steps 6 --batch 32
run emo_n2 900 $TR scripts/emotion_step_bench.py --autocast --steps 6
