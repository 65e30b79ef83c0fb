#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-1500; }
TAILN=12 run t_frontend 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py tests/test_gpu_emotion.py -q -m gpu -x
TAILN=5 run l0bench 600 python scripts/bench_layer0.py
run bench 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["clocks"])
for k in d["kernels"]: print(k["kernel"][:50], round(k["ms"],4), round(k["achieved"]), round(k["frac"],3))
PY
