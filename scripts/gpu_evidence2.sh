#!/bin/bash
# Evidence for profiles/ (second half of round 1): default bench line, ncu launch lists of the same commands, full ncu
# captures of the new kernels (resident mix, optimizer tail, attentive pooling), optimizer / step A-B at batch 64.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1200; }
run bench_full 900 python bench.py
BENCH="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
run bench_plain 600 $BENCH
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
PK="python scripts/profile_kernels.py"
run pk_plain 600 $PK
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pk.csv $PK > gpurun_out/ncu1b.log 2>&1; echo "pk launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mix_normalize_resident|grad_sqnorm|adamw_ema|asp_" -s 10 -c 10 -f -o gpurun_out/prof_new $PK > gpurun_out/ncu2.log 2>&1; echo "new-kernel capture rc=$?"
run optimbench 600 python scripts/bench_optim.py
run mixbench 600 python scripts/bench_mix.py
run byol64_fused 900 python scripts/train_step_bench.py --autocast --steps 10 --batch 64 --layerdrop 0
run byol64_torch 900 python scripts/train_step_bench.py --autocast --steps 10 --batch 64 --layerdrop 0 --optimizer torch
run emo_fused 900 python scripts/emotion_step_bench.py --autocast --steps 10 --layerdrop 0
run emo_torch 900 python scripts/emotion_step_bench.py --autocast --steps 10 --layerdrop 0 --optimizer torch
run emo_stock 900 python scripts/emotion_step_bench.py --autocast --steps 10 --layerdrop 0 --optimizer torch --stock-pool
ls -la gpurun_out | head -40
