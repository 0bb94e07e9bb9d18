#!/bin/bash
# Round-2 second measurement batch: GEMM epilogue diet (fast divmod, 256-thread packed-fp32x2 statistics, residual prefetch
# one tile ahead), deeper load queues in the BN kernels.
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=900 TAIL=40 run python -m pytest tests -m gpu -q -x
T=500 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 20 --warmup 5 --dump-gemm-profile gpurun_out/r02b_gemm_launches.json
T=600 TAIL=3 run ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_step.csv python scripts/profile_step.py
T=300 TAIL=3 run ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 5 -o gpurun_out/r02b_gemm_cases python scripts/ncu_gemm_cases.py l1conv3 l1dgrad l1conv l2wgrad l3conv
