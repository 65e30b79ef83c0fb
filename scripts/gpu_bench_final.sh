#!/bin/bash
# final bench lines of the round: the default command (200 steps) and the 20-step command the launch list belongs to
mkdir -p gpurun_out
T=${TAG:-r2z}
NRSE_BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/${T}_bench_full.log 2> gpurun_out/${T}_bench_full.err; echo "full rc=$?"
NRSE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_20.log 2> gpurun_out/${T}_bench_20.err; echo "20 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.log 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
for f in full 20 ref; do tail -1 gpurun_out/${T}_bench_$f.log | cut -c1-330; done
grep -h "bench debug" gpurun_out/${T}_bench_full.err gpurun_out/${T}_bench_20.err | cut -c1-160
