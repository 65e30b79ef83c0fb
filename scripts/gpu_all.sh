#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-1200; }
TAILN=6 run t_all 1800 python -m pytest tests -q -m gpu
run smoke 600 python __graft_entry__.py smoke
TAILN=5 run l0bench 600 python scripts/bench_layer0.py
run poolbench 600 python scripts/bench_pool.py
run bench 900 python bench.py
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["clocks"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
for k in d["kernels"]: print(k["kernel"][:50], round(k["ms"],4), round(k["achieved"]), round(k["frac"],3))
print(d["stock_torch_same_gpu_ms"]); print(d["frontend_train"])
PY
