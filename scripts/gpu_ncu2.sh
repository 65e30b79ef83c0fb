#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/profile_kernels.py"
$CMD > gpurun_out/ncu_plain2.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/ncu_plain2.log; exit 1; }
# skip the first iteration (warm-up): 6 pack + 2 mix + 7 frontend = 15 launches
ncu --set full --clock-control none --import-source on -s 21 -c 9 -f -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_b.log 2>&1
echo "capture rc=$?"; tail -3 gpurun_out/ncu_b.log
