#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) one full capture of the
# dominant kernel (tcgen05 conv GEMM), (3) full capture of the bandwidth-bound kernels.  Plain run first.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 12 -c 3 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu2.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"mix_normalize|layer0_kernel|ema_chunks|byol_loss" -s 6 -c 6 -f -o gpurun_out/prof_bw $CMD > gpurun_out/ncu3.log 2>&1
echo "bw capture rc=$?"
ls -la gpurun_out/
