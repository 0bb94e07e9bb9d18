#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=400 TAIL=30 run python -m pytest tests/test_gpu_kernels.py -m gpu -q
T=500 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 20 --warmup 5 --dump-gemm-profile gpurun_out/r02g_gemm_launches.json
T=900 TAIL=30 run python -m pytest tests/test_gpu_parity.py -m gpu -q
T=600 TAIL=200 run python scripts/library_ab.py profiles/r02f_gemm_launches.json
