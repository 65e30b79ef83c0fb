#!/bin/bash
# Timing experiments with deliberately wrong results (NRSE_EXPERIMENT: 1 = epilogues store nothing, 2 = GEMM layers skip the
# LayerNorm statistics pass, 3 = both) next to the real kernels: where does the epilogue's time go?
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
TAILN=6 run exp_tests 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py -q -m gpu -x
for e in 0 1 2 3; do
  echo "=== NRSE_EXPERIMENT=$e"
  NRSE_EXPERIMENT=$e TAILN=2 run exp_gemm_$e 300 python scripts/bench_gemm.py
  NRSE_EXPERIMENT=$e TAILN=4 run exp_l0_$e 300 python scripts/bench_layer0.py
done
