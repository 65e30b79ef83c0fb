#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-1500; }
TAILN=12 run t_new 900 python -m pytest tests/test_gpu_emotion.py tests/test_gpu_optim.py tests/test_gpu_mix.py -q -m gpu
TAILN=10 run stagger 600 python scripts/bench_mix_stagger.py
run emo_fused 900 python scripts/emotion_step_bench.py --autocast --steps 5
run byol_fused 900 python scripts/train_step_bench.py --autocast --steps 5 --batch 32
run byol_torch 900 python scripts/train_step_bench.py --autocast --steps 5 --batch 32 --optimizer torch
