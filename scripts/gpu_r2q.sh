#!/bin/bash
# round 2, re-entry pass: full GPU test suite, smoke(), kernel timeline of the headline step (eager / graph), one bench run (N=1)
mkdir -p gpurun_out
T=${TAG:-r2q}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 300 python scripts/step_timeline.py --graph > gpurun_out/${T}_step_timeline.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; tail -4 gpurun_out/${T}_step_timeline.log; tail -c 600 gpurun_out/${T}_bench.err
