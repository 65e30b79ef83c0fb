#!/bin/bash
# N-GPU tuning sweep of the training-step leg: SMs reserved for NCCL x NCCL CTA caps x bucket size
N=${1:-2}
mkdir -p gpurun_out
out=gpurun_out/${TAG:-r2}_train_tune_n$N.jsonl
: > $out
run() {  # reserve bucket_mb [env...]
  local reserve=$1 bucket=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 30 --warmup 3 --train-only --train-sm-reserve $reserve --train-bucket-mb $bucket 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read())['train_step']; print(json.dumps({'cfg':'$reserve $bucket $*','ms':d['ms_per_step'],'nosync':d['ms_per_step_no_allreduce'],'exposed':d['allreduce_exposed_ms']}))" >> $out
}
run 0 256 X=1
run 8 256 X=1
run 16 256 X=1
run 32 256 X=1
run 16 64 X=1
run 16 256 NCCL_MAX_CTAS=8
run 16 256 NCCL_MAX_CTAS=16
run 32 256 NCCL_MAX_CTAS=32
run 0 256 NCCL_MAX_CTAS=8
cat $out
