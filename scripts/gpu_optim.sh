#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-1500; }
TAILN=25 run t_optim 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_mix.py -q -m gpu -x
run optimbench 600 python scripts/bench_optim.py
TAILN=12 run mixbench 600 python scripts/bench_mix.py
run bench 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
for k in d["kernels"]: print(k["kernel"][:50], round(k["ms"],4), round(k["achieved"]), round(k["frac"],3))
print(d.get("stock_torch_same_gpu_ms"))
PY
