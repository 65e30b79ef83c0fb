#!/bin/bash
# A/B of the fused data gradient + LayerNorm / GELU backward; optional ncu capture of the fused kernel (NCU=1)
mkdir -p gpurun_out
T=${TAG:-r2}
timeout 300 python -m pytest tests/test_gpu_frontend_bwd.py -q --timeout 120 -p no:cacheprovider -x -k "fused or full_backward_vs" 2>&1 | tail -8
timeout 300 python scripts/bench_dgrad_fused.py 1 2 3 4 > gpurun_out/${T}_dgf.jsonl 2> gpurun_out/${T}_dgf.err; cat gpurun_out/${T}_dgf.jsonl; tail -3 gpurun_out/${T}_dgf.err
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgrad_lnbwd -o gpurun_out/${T}_dgf_ncu -f \
    python scripts/bench_dgrad_fused.py 1 --once > gpurun_out/${T}_dgf_ncu.log 2>&1; tail -2 gpurun_out/${T}_dgf_ncu.log
fi
