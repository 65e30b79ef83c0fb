#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-1500; }
TAILN=8 run t_emotion 900 python -m pytest tests/test_gpu_emotion.py -q -m gpu
run poolbench 600 python scripts/bench_pool.py
PK="python scripts/profile_kernels.py"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mix_normalize_resident|asp_" -c 12 -f -o gpurun_out/prof_new2 $PK > gpurun_out/ncu3.log 2>&1; echo "capture rc=$?"
