#!/bin/bash
# kernel timeline (CUPTI through torch.profiler) of the backward, separate and fused form
mkdir -p gpurun_out
T=${TAG:-r2}
timeout 600 python scripts/bench_bwd.py --timeline > gpurun_out/${T}_bwd_timeline.log 2>&1; tail -3 gpurun_out/${T}_bwd_timeline.log
