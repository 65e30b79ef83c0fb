#!/bin/bash
# round 2, second GPU pass (feature projection, tolerance table, bench fixes): full GPU test suite, smoke(), one bench run (N=1)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG:-r2}_gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG:-r2}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG:-r2}_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG:-r2}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${TAG:-r2}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG:-r2}_bench.log 2> gpurun_out/${TAG:-r2}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG:-r2}_bench.err
tail -3 gpurun_out/${TAG:-r2}_tests.log; tail -2 gpurun_out/${TAG:-r2}_smoke.log; tail -c 600 gpurun_out/${TAG:-r2}_bench.err
