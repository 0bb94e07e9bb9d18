#!/bin/bash
# 2-GPU data-parallel check: bench line with dp_check (post-all-reduce gradients == mean of per-rank gradients on NCCL)
set -u
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu 2>&1 | tail -3 | tee gpurun_out/r02_bench_n2.txt
