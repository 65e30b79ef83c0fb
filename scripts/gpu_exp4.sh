#!/bin/bash
# SM clock inside the layer-1 GEMM launches (NRSE_EXPERIMENT flag 256: clock64 / globaltimer printf at kernel end).  The hook
# was removed from the kernels after the measurement (profiles/r1c_epilogue_experiments.txt); check out commit 5ed733e to re-run.
mkdir -p gpurun_out
for e in 256 257 260 264; do
  echo "=== NRSE_EXPERIMENT=$e"
  NRSE_EXPERIMENT=$e timeout 300 python scripts/bench_gemm.py > gpurun_out/exp4_gemm_$e.log 2>&1
  grep "tiles=3200" gpurun_out/exp4_gemm_$e.log | sort | uniq -c | sort -rn | head -6
  grep "variant=" gpurun_out/exp4_gemm_$e.log | head -2 | cut -c1-200
done
