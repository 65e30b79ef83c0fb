#!/bin/bash
# Final evidence of the round: full GPU test suite, smoke, default bench line (+ ncu launch list of the same command),
# micro-benchmarks, configs[2..4] scripts.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-700; }
TAILN=4 run t_all 1800 python -m pytest tests -q -m gpu
run smoke 600 python __graft_entry__.py smoke
run bench_full 900 python bench.py
BENCH="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
run bench_plain 600 $BENCH
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
TAILN=5 run l0bench 300 python scripts/bench_layer0.py
TAILN=4 run gemmbench 300 python scripts/bench_gemm.py
TAILN=4 run frontendbench 300 python scripts/bench_frontend.py
run bwdbench 600 python scripts/bench_bwd.py
run optimbench 600 python scripts/bench_optim.py
run poolbench 300 python scripts/bench_pool.py
TAILN=20 run mixbench 600 python scripts/bench_mix.py
TAILN=8 run evalsweep 600 python scripts/eval_sweep_bench.py
run byol64 900 python scripts/train_step_bench.py --autocast --steps 10 --batch 64 --layerdrop 0 --profile
run emo 900 python scripts/emotion_step_bench.py --autocast --steps 10 --layerdrop 0
