#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -3
timeout 300 python scripts/bench_bwd.py 2>&1 | tail -1 | cut -c1-400
