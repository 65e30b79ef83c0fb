#!/bin/bash
# ncu launch list (per-launch durations) of scripts/profile_kernels.py
mkdir -p gpurun_out
timeout 600 python scripts/profile_kernels.py > gpurun_out/${TAG:-r2}_pk_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-r2}_launches_pk.csv \
  python scripts/profile_kernels.py > gpurun_out/${TAG:-r2}_pk_ncu.log 2>&1
tail -2 gpurun_out/${TAG:-r2}_pk_ncu.log
