#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
run t_mix 900 python -m pytest tests/test_gpu_mix.py -q -m gpu -x
TAILN=80 run mixbench 600 python scripts/bench_mix.py
