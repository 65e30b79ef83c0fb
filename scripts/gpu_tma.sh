#!/bin/bash
# Parity tests of the conv frontend (forward + backward) and the micro-benchmarks, previous build vs current build.
mkdir -p gpurun_out
CS=noise-robust-speech-embedding_b200/csrc
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
TAILN=8 run tma_tests 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py -q -m gpu
for lib in ${LIBS:-default prev}; do
  if [ "$lib" = default ]; then unset NRSE_B200_LIB; else export NRSE_B200_LIB=$PWD/$CS/build/libnrse_b200_$lib.so; fi
  echo "=== build: $lib"
  TAILN=4 run tma_gemm_$lib 300 python scripts/bench_gemm.py
  TAILN=4 run tma_l0_$lib 300 python scripts/bench_layer0.py
  TAILN=2 run tma_front_$lib 300 python scripts/bench_frontend.py
done
unset NRSE_B200_LIB
TAILN=3 run tma_bwd 600 python scripts/bench_bwd.py
