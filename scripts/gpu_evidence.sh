#!/bin/bash
# Evidence for profiles/: plain bench, ncu launch list of the SAME bench command, full ncu captures of every kernel.
mkdir -p gpurun_out
BENCH="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
$BENCH > gpurun_out/bench_plain.log 2>&1 || { echo "plain bench failed"; tail -20 gpurun_out/bench_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
PK="python scripts/profile_kernels.py"
$PK > gpurun_out/pk_plain.log 2>&1 || { echo "plain profile_kernels failed"; tail -20 gpurun_out/pk_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pk.csv $PK > gpurun_out/ncu1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mix_normalize|layer0_tc|conv_gemm" -s 22 -c 9 -f -o gpurun_out/prof_fwd $PK > gpurun_out/ncu2.log 2>&1
echo "fwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ln_gelu_bwd|conv_wgrad|layer0_wgrad" -s 14 -c 14 -f -o gpurun_out/prof_bwd $PK > gpurun_out/ncu3.log 2>&1
echo "bwd capture rc=$?"
ls -la gpurun_out | head -30
