#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
TAILN=4 run d_tests 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py tests/test_gpu_model.py -q -m gpu
run d_bwd1 600 python scripts/bench_bwd.py
run d_bwd2 600 python scripts/bench_bwd.py
