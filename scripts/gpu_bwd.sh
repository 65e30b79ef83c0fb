#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-900; }
TAILN=5 run t_bwd 900 python -m pytest tests/test_gpu_frontend_bwd.py tests/test_gpu_frontend.py tests/test_gpu_model.py -q -m gpu
run bwdbench 600 python scripts/bench_bwd.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bwd.csv python scripts/profile_kernels.py > gpurun_out/ncu_bwd.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_bwd.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hdr])}
from collections import defaultdict
agg=defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)<len(col) or r[col['Metric Name']]!='gpu__time_duration.sum': continue
    n=r[col['Kernel Name']][:50]; v=float(r[col['Metric Value']].replace(',','')); u=r[col['Metric Unit']]
    v = v/1000 if u in ('ns','nsecond') else v
    agg[n][0]+=1; agg[n][1]+=v
for n,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print(f"{c:5d} launches {t/c:9.1f} us avg {t:10.1f} us total  {n}")
PY
