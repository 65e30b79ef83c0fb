#!/bin/bash
# programmatic dependent launch between the frontend kernels: parity, then A/B (NRSE_EXPERIMENT=512 switches it off)
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=3 run pdl_tests 1200 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py tests/test_gpu_model.py -q -m gpu
for e in 0 512 0 512; do
  echo "== NRSE_EXPERIMENT=$e"
  NRSE_EXPERIMENT=$e timeout 300 python scripts/bench_frontend.py 2>&1 | grep "tile order 1" | tail -1
  NRSE_EXPERIMENT=$e timeout 300 python scripts/eval_sweep_bench.py 2>&1 | grep '"seconds": 4' | cut -c1-120
done
run pdl_bench 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/pdl_bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
PY
