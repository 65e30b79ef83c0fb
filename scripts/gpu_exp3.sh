#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
for e in ${EXPS:-0 8 16 24 56 64 65 129}; do
  echo "=== NRSE_EXPERIMENT=$e"
  NRSE_EXPERIMENT=$e TAILN=2 run exp3_gemm_$e 300 python scripts/bench_gemm.py
  NRSE_EXPERIMENT=$e TAILN=4 run exp3_l0_$e 300 python scripts/bench_layer0.py
done
