#!/bin/bash
# 2-GPU experiment: does capping NCCL's CTA count reduce the interference with the persistent GEMM?
set -u
mkdir -p gpurun_out
for cfg in "" "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=8" "NCCL_MAX_CTAS=16" "NCCL_MAX_CTAS=8 NCCL_ALGO=Ring"; do
  echo "=== $cfg"
  env $cfg timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu 2>&1 | grep '^{"metric"' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['dp_check'])"
done
