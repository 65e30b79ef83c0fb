#!/bin/bash
# quick loop: frontend / model parity, then a short bench line
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
TAILN=8 run q_tests 1200 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_frontend_bwd.py tests/test_gpu_model.py -q -m gpu
TAILN=3 run q_gemm 300 python scripts/bench_gemm.py
TAILN=4 run q_l0 300 python scripts/bench_layer0.py
TAILN=2 run q_front 300 python scripts/bench_frontend.py
run q_bench 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
l=[x for x in open('gpurun_out/q_bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d.get("clocks"))
print("gemm", d["roofline"]["achieved"], d["roofline"]["frac"], [ (p["layer"], round(p["ms"],4), round(p["tflops"])) for p in d["roofline"]["per_layer"]])
for k in d["kernels"]: print(k["kernel"][:50], round(k["ms"],4), round(k["achieved"]), round(k["frac"],3))
PY
