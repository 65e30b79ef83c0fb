#!/bin/bash
# the short (20-step) bench line several times in fresh processes: is the first timed window stable?  WARMUP=n sets --warmup
mkdir -p gpurun_out
T=${TAG:-r2x}
W=${WARMUP:-5}
for i in 1 2 3 4; do
  NRSE_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup $W --no-train-step --no-cpu-baseline --sustain-seconds 0.3 > gpurun_out/${T}_rep$i.log 2> gpurun_out/${T}_rep$i.err
  python - <<PY
import json
l=[x for x in open('gpurun_out/${T}_rep$i.log') if x.startswith('{"metric')][-1]
d=json.loads(l)
print('run $i warmup $W value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sus', round(d['sustained']['ms_per_step'],3))
PY
  grep "bench debug" gpurun_out/${T}_rep$i.err | head -1
done
