#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
TAILN=2 run write_bw 300 python scripts/bench_write_bw.py
for e in 0 4 1; do
  echo "=== NRSE_EXPERIMENT=$e"
  NRSE_EXPERIMENT=$e TAILN=2 run exp2_gemm_$e 300 python scripts/bench_gemm.py
  NRSE_EXPERIMENT=$e TAILN=4 run exp2_l0_$e 300 python scripts/bench_layer0.py
done
