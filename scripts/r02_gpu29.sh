#!/bin/bash
# compute-sanitizer over the fused-BN-reduce / CTA-pair GEMM tests (memcheck: out-of-bounds global / shared accesses;
# racecheck: shared-memory hazards of the epilogue)
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=280 TAIL=40 run compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "over_residual_is_exact and 256"
T=200 TAIL=40 run compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cta_pairs_match and 1000"
T=280 TAIL=60 run compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "over_residual_is_exact and 256"
