#!/bin/bash
# Evidence for profiles/ (third part of round 1, r1c_*): full GPU test suite, smoke, the default bench line and the ncu
# launch list of the same command, a full ncu capture of the forward kernels after the epilogue rework (staged TMA output
# stores, evict-first policy, 16-warp layer 0, 2-SM UMMA kernel on layers 1-2), micro-benchmarks.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-900; }
TAILN=4 run c_tests 1800 python -m pytest tests -q -m gpu
run c_smoke 600 python __graft_entry__.py smoke
run c_bench_full 900 python bench.py
BENCH="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
run c_bench_plain 600 $BENCH
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c_launches.csv $BENCH > gpurun_out/c_ncu1.log 2>&1; echo "launch list rc=$?"
PK="python scripts/profile_kernels.py"
run c_pk_plain 600 $PK
# the inference forward of the driver: layer 0 + six GEMM layers
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"layer0_tc_kernel|conv_gemm" -s 0 -c 7 -f -o gpurun_out/prof_r1c $PK > gpurun_out/c_ncu2.log 2>&1; echo "forward capture rc=$?"
TAILN=4 run c_gemm 300 python scripts/bench_gemm.py
TAILN=4 run c_l0 300 python scripts/bench_layer0.py
TAILN=4 run c_front 300 python scripts/bench_frontend.py
run c_bwd 600 python scripts/bench_bwd.py
TAILN=8 run c_evalsweep 600 python scripts/eval_sweep_bench.py
run c_byol64 900 python scripts/train_step_bench.py --autocast --steps 10 --batch 64 --layerdrop 0
run c_emo 900 python scripts/emotion_step_bench.py --autocast --steps 10 --layerdrop 0
