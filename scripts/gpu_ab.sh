#!/bin/bash
# A/B of libnrse_b200 builds on ONE box (NRSE_B200_LIB picks the build): frontend parity tests on the default build, then
# the GEMM-layer / layer-0 / whole-frontend micro-benchmarks for every build listed in $LIBS (names under csrc/build/).
mkdir -p gpurun_out
CS=noise-robust-speech-embedding_b200/csrc
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
TAILN=6 run ab_tests 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py -q -m gpu -x
for lib in ${LIBS:-prev default}; do
  if [ "$lib" = default ]; then unset NRSE_B200_LIB; else export NRSE_B200_LIB=$PWD/$CS/build/libnrse_b200_$lib.so; fi
  echo "=== build: $lib"
  TAILN=4 run ab_gemm_$lib 300 python scripts/bench_gemm.py
  TAILN=4 run ab_l0_$lib 300 python scripts/bench_layer0.py
  TAILN=2 run ab_front_$lib 300 python scripts/bench_frontend.py
done
unset NRSE_B200_LIB
