#!/bin/bash
# round 2, first GPU pass: full GPU test suite, smoke(), one bench run (N=1)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_tests.log; tail -2 gpurun_out/r2a_smoke.log; tail -c 600 gpurun_out/r2a_bench.err
