#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-1500; }
TAILN=5 run t_new 900 python -m pytest tests/test_gpu_emotion.py tests/test_gpu_optim.py tests/test_gpu_mix.py tests/test_gpu_model.py -q -m gpu
run emo_fused 900 python scripts/emotion_step_bench.py --autocast --steps 8
run emo_fused_keep 900 python scripts/emotion_step_bench.py --autocast --steps 8 --keep-grads
run emo_torch 900 python scripts/emotion_step_bench.py --autocast --steps 8 --optimizer torch
run emo_stock 900 python scripts/emotion_step_bench.py --autocast --steps 8 --optimizer torch --stock-pool
run byol_fused 900 python scripts/train_step_bench.py --autocast --steps 8 --batch 32
run byol_fused_keep 900 python scripts/train_step_bench.py --autocast --steps 8 --batch 32 --keep-grads
run byol_torch 900 python scripts/train_step_bench.py --autocast --steps 8 --batch 32 --optimizer torch
run bench1 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline
run bench2 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --two-streams
python - <<'PY'
import json
for f in ("bench1","bench2"):
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{')][-1]
    d=json.loads(l)
    print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["clocks"])
PY
