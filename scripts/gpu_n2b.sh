#!/bin/bash
# 2-GPU evidence after the r1c kernel work: hot-path bench under torchrun (default steps) and the reference arm
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-2} gpurun_out/$name.log | cut -c1-700; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run n2_bench 900 $TR bench.py --gpus 2 --steps 100 --warmup 3
run n2_bench_ref 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1
run n2_tests 900 python -m pytest tests/test_gpu_model.py -q -m gpu
