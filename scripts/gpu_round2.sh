#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 12 gpurun_out/$name.log; }
run t_mix 400 python -m pytest tests/test_gpu_mix.py -q -m gpu -x
run t_layer0 300 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k "layer0"
run t_frontend 500 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k "not layer0"
run t_rest 300 python -m pytest tests/test_gpu_ema_loss.py -q -m gpu
run bench 600 python bench.py --steps 50 --warmup 3
