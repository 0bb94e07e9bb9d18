#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=900 TAIL=80 run python scripts/tune_gemm.py profiles/r02l_gemm_launches_dual_all.json --top 70
