#!/bin/bash
# first GPU bring-up: every stage under its own timeout so that one hang does not hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 15 gpurun_out/$name.log; }
run t_mix_ema_loss 400 python -m pytest tests/test_gpu_mix.py tests/test_gpu_ema_loss.py -q -m gpu
run t_layer0 200 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k "layer0"
run t_gemm_v1 200 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k "gemm_layer and v1"
run t_gemm_v2 200 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k "gemm_layer and v2"
run t_frontend 400 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k "not gemm_layer and not layer0"
run smoke 200 python __graft_entry__.py smoke
run bench 600 python bench.py --steps 10 --warmup 3
