#!/bin/bash
# full ncu capture of the 2-SM UMMA kernel on layer 1 (first conv_gemm2_kernel launch of the driver)
mkdir -p gpurun_out
PK="python scripts/profile_kernels.py"
timeout 600 $PK > gpurun_out/g2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/g2_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm2" -s 0 -c 1 -f -o gpurun_out/prof_gemm2 $PK > gpurun_out/g2_ncu.log 2>&1; echo "capture rc=$?"
