#!/bin/bash
# 4 GPUs: CTA pairs on / off under the overlapped NCCL gradient exchange
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
D="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 30 --warmup 5 --skip-cpu"
T=400 TAIL=1 run $D
T=400 TAIL=1 run env VTX_GEMM_PAIR=0 $D
