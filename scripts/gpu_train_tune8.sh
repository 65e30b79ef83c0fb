#!/bin/bash
# 8-GPU: all-reduce variants of the training-step leg
N=${1:-8}
mkdir -p gpurun_out
out=gpurun_out/${TAG:-r2}_train_tune2_n$N.jsonl
: > $out
run() {  # impl reserve bucket_mb ctas [env...]
  local impl=$1 reserve=$2 bucket=$3 ctas=$4; shift 4
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 30 --warmup 3 --train-only --train-allreduce-impl $impl --train-sm-reserve $reserve --train-bucket-mb $bucket --train-multimem-ctas $ctas 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read())['train_step']; print(json.dumps({'cfg':'$impl reserve=$reserve bucket=$bucket ctas=$ctas $*','ms':d['ms_per_step'],'nosync':d['ms_per_step_no_allreduce'],'exposed':d['allreduce_exposed_ms']}))" >> $out
}
run nccl 32 2048 0 X=1
run nccl 32 2048 0 NCCL_ALGO=NVLS
run nccl 32 2048 0 NCCL_ALGO=Ring
run multimem 0 256 64 X=1
run multimem 0 256 148 X=1
run multimem 32 256 32 X=1
cat $out
