#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=300 TAIL=80 run python scripts/debug/diag_engine_bnr.py
T=600 TAIL=40 run python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cta_pairs"
