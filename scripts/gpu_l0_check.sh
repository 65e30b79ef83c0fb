#!/bin/bash
mkdir -p gpurun_out
CS=noise-robust-speech-embedding_b200/csrc
for lib in ${LIBS:-default committed}; do
  if [ "$lib" = default ]; then unset NRSE_B200_LIB; else export NRSE_B200_LIB=$PWD/$CS/build/libnrse_b200_$lib.so; fi
  echo "== $lib"; timeout 300 python scripts/bench_layer0.py 2>&1 | grep "variant [23]"
done
