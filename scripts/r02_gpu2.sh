#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=600 TAIL=40 run python -m pytest tests -m gpu -q
T=300 TAIL=30 run python scripts/debug/dbg_bias_h2048.py
T=400 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3 --dump-gemm-profile gpurun_out/r02_gemm_launches_a.json
