#!/bin/bash
# r1c evidence for the kernels around the inference forward: ncu launch list of scripts/profile_kernels.py (every hot-path kernel:
# mix, forward, tape-writing forward, backward, optimizer tail, pooling) and a full capture of one backward pass.
mkdir -p gpurun_out
PK="python scripts/profile_kernels.py"
timeout 600 $PK > gpurun_out/d_pk_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/d_pk_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_pk.csv $PK > gpurun_out/d_ncu1.log 2>&1; echo "launch list rc=$?"
# inference forward: 3 conv_gemm_kernel launches (layers 4-6), tape-writing forward: 6 -> skip 9; one backward = 7 ln_gelu_bwd +
# 6 conv_wgrad + 12 data-gradient conv_gemm_kernel + 1 layer0_wgrad
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ln_gelu_bwd|conv_wgrad|layer0_wgrad|conv_gemm_kernel" -s 9 -c 26 -f -o gpurun_out/prof_r1c_bwd $PK > gpurun_out/d_ncu2.log 2>&1; echo "backward capture rc=$?"
