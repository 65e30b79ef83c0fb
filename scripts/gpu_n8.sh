#!/bin/bash
# 8-GPU pass: NCCL / backward overlap probe, then the training-step leg at a few SM reservations
N=${1:-8}
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  scripts/nccl_overlap_probe.py > gpurun_out/${TAG:-r2}_probe_n$N.log 2>&1
grep -v "NCCL INFO" gpurun_out/${TAG:-r2}_probe_n$N.log | tail -1
grep -o "Algo [A-Za-z_]* proto [A-Za-z0-9_]*\|algorithm[^,]*\|NVLS[^,]*comm[^,]*" gpurun_out/${TAG:-r2}_probe_n$N.log | sort | uniq -c | head
out=gpurun_out/${TAG:-r2}_train_tune_n$N.jsonl
: > $out
for reserve in 0 16 32; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 30 --warmup 3 --train-only --train-sm-reserve $reserve 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read())['train_step']; print(json.dumps({'reserve':$reserve,'ms':d['ms_per_step'],'nosync':d['ms_per_step_no_allreduce'],'exposed':d['allreduce_exposed_ms'],'value':d['value']}))" >> $out
done
cat $out
