#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend_bwd.py -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -15
timeout 600 python scripts/bench_bwd.py --quick > gpurun_out/${TAG:-r2}_bwd.log 2>&1; tail -2 gpurun_out/${TAG:-r2}_bwd.log
