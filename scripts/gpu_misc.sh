#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_tail_modules.py > gpurun_out/${TAG:-r2}_tail_modules.log 2>&1; tail -1 gpurun_out/${TAG:-r2}_tail_modules.log
timeout 600 python scripts/bench_bwd.py --quick > gpurun_out/${TAG:-r2}_bwd.log 2>&1; tail -1 gpurun_out/${TAG:-r2}_bwd.log
