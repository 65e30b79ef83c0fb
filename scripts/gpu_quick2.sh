#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 6 gpurun_out/$name.log | cut -c1-600; }
run t_mix 600 python -m pytest tests/test_gpu_mix.py tests/test_gpu_model.py -q -m gpu -x
run bench 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("gemm", d["roofline"]["achieved"], [ (p["layer"], round(p["ms"],4), round(p["tflops"])) for p in d["roofline"]["per_layer"]])
for k in d["kernels"]: print(k["kernel"][:50], round(k["ms"],4), round(k["achieved"]), round(k["frac"],3))
PY
