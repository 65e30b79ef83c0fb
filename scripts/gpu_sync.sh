#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-900; }
TAILN=6 run t_mixmodel 900 python -m pytest tests/test_gpu_model.py -q -m gpu
run byol64_fused 900 python scripts/train_step_bench.py --autocast --steps 10 --batch 64 --layerdrop 0
run byol64_torch 900 python scripts/train_step_bench.py --autocast --steps 10 --batch 64 --layerdrop 0 --optimizer torch
run emo_fused 900 python scripts/emotion_step_bench.py --autocast --steps 10 --layerdrop 0
