#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layer0_tc_kernel" -s 45 -c 1 -f -o gpurun_out/prof_l0_v2 python scripts/bench_layer0.py > gpurun_out/ncu_l0b.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_l0b.log
