#!/bin/bash
# Round-2 third measurement batch: three-buffer residual pipeline, scalar 256-thread BN statistics, grouped dropout hash,
# direct forward max-pool; per-class tile_n / split_k sweep; 3 bench configs for the record.
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=900 TAIL=40 run python -m pytest tests -m gpu -q
T=500 TAIL=1 run python bench.py --steps 20 --warmup 5 --dump-gemm-profile gpurun_out/r02c_gemm_launches.json
T=600 TAIL=3 run ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_step.csv python scripts/profile_step.py
T=900 TAIL=80 run python scripts/tune_gemm.py profiles/r02a_gemm_launches.json --top 70
T=300 TAIL=3 run ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 5 -o gpurun_out/r02c_gemm_cases python scripts/ncu_gemm_cases.py l1conv3 l1dgrad l1conv l2wgrad l3conv
T=300 TAIL=3 run ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_|bn_bwd" -c 10 -o gpurun_out/r02c_attn_bn python scripts/profile_step.py
