#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-700; }
TAILN=6 run t_front 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py tests/test_gpu_model.py -q -m gpu
TAILN=5 run l0bench 300 python scripts/bench_layer0.py
TAILN=4 run gemmbench 300 python scripts/bench_gemm.py
run bench 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["clocks"], d.get("kernel_clocks"))
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
for k in d["kernels"]: print(k["kernel"][:50], round(k["ms"],4), round(k["achieved"]), round(k["frac"],3))
PY
