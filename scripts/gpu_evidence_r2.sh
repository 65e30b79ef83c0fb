#!/bin/bash
# Evidence for profiles/ (round 2, final state, r2z_*): full GPU test suite, smoke, the default bench line (200 steps), the
# 20-step line and the ncu launch list of that same command, a full ncu capture of the inference forward kernels (feeds
# roofline.traffic), the headline step's kernel timeline, the backward micro-benchmark.
mkdir -p gpurun_out
T=${TAG:-r2z}
run() { name=$1; shift; timeout "$@" > gpurun_out/${T}_$name.log 2> gpurun_out/${T}_$name.err; echo "$name rc=$?"; tail -n ${TAILN:-2} gpurun_out/${T}_$name.log | cut -c1-700; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
run smoke 600 python __graft_entry__.py smoke
run bench_full 900 python bench.py
BENCH="python bench.py --steps 20 --warmup 5"
run bench_20 600 $BENCH
run step_timeline 300 python scripts/step_timeline.py
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv $BENCH --no-cpu-baseline > gpurun_out/${T}_ncu1.log 2>&1; echo "launch list rc=$?"
PK="python scripts/profile_kernels.py"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"layer0_tc_kernel|conv_gemm|mix_normalize_resident" -s 0 -c 9 -f -o gpurun_out/${T}_prof_fwd $PK > gpurun_out/${T}_ncu2.log 2>&1; echo "forward capture rc=$?"
run bwd 600 python scripts/bench_bwd.py --quick
