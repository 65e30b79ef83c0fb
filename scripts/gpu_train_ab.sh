#!/bin/bash
mkdir -p gpurun_out
CS=noise-robust-speech-embedding_b200/csrc
for lib in prev committed default prev committed default; do
  if [ "$lib" = default ]; then unset NRSE_B200_LIB; else export NRSE_B200_LIB=$PWD/$CS/build/libnrse_b200_$lib.so; fi
  timeout 600 python scripts/bench_bwd.py > gpurun_out/e_bwd_$lib.log 2>&1
  echo "$lib: $(tail -1 gpurun_out/e_bwd_$lib.log | cut -c1-140)"
done
