#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=400 TAIL=30 run python -m pytest tests/test_gpu_kernels.py -m gpu -q
T=500 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5 --dump-gemm-profile gpurun_out/r02j_gemm_launches.json
T=900 TAIL=30 run python -m pytest tests/test_gpu_parity.py -m gpu -q
T=600 TAIL=3 run ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02j_launches_step.csv python scripts/profile_step.py
T=500 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5
