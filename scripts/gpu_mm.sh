#!/bin/bash
# N-GPU: single-GPU tests of the new kernels, multicast all-reduce check, then the training-step leg with both all-reduce paths
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_posconv.py tests/test_gpu_frontend_bwd.py tests/test_gpu_featproj.py -q --timeout 300 -p no:cacheprovider > gpurun_out/${TAG:-r2}_tests_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG:-r2}_tests_new.log
tail -5 gpurun_out/${TAG:-r2}_tests_new.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/multimem_check.py > gpurun_out/${TAG:-r2}_mmcheck_n$N.log 2>&1
echo "mmcheck rc=$?"; tail -3 gpurun_out/${TAG:-r2}_mmcheck_n$N.log | cut -c1-600
out=gpurun_out/${TAG:-r2}_train_mm_n$N.jsonl
: > $out
for cfg in "nccl 32 0" "multimem 0 0" "multimem 0 64" "multimem 16 0" "multimem 0 32"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 30 --warmup 3 --train-only --train-allreduce-impl $1 --train-sm-reserve $2 --train-multimem-ctas $3 2>gpurun_out/${TAG:-r2}_train_mm.err | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read())['train_step']; print(json.dumps({'cfg':'$cfg','ms':d['ms_per_step'],'nosync':d['ms_per_step_no_allreduce'],'exposed':d['allreduce_exposed_ms'],'impl':d['allreduce_impl'][:40],'loss':d['loss']}))" >> $out
done
cat $out; tail -5 gpurun_out/${TAG:-r2}_train_mm.err | cut -c1-300
