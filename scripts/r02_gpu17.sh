#!/bin/bash
# 2-GPU box: GPU tests (fused BN-backward reduction, both GEMM schedules), N=1 A/Bs, N=2 static vs dynamic schedule
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
export CUDA_VISIBLE_DEVICES=0
T=1500 TAIL=15 run python -m pytest tests -m gpu -x -q
B="python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5"
T=400 TAIL=1 run $B --dump-gemm-profile gpurun_out/r02n_gemm_launches_bnr.json
T=400 TAIL=1 run env VTX_BNR_FUSE=0 $B --dump-gemm-profile gpurun_out/r02n_gemm_launches_nobnr.json
T=400 TAIL=1 run env VTX_GEMM_SCHEDULE=dynamic $B --dump-gemm-profile gpurun_out/r02n_gemm_launches_bnr_dynamic.json
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_BNR_FUSE=0 $B
unset CUDA_VISIBLE_DEVICES
D="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu"
T=500 TAIL=1 run env VTX_GEMM_SCHEDULE=static $D
T=500 TAIL=1 run env VTX_GEMM_SCHEDULE=dynamic $D
T=500 TAIL=1 run env VTX_GEMM_SCHEDULE=static $D
T=500 TAIL=1 run env VTX_GEMM_SCHEDULE=dynamic $D
