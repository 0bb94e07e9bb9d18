#!/bin/bash
# compute-sanitizer racecheck over more GEMM epilogue configurations (CTA pairs, two-group epilogue with statistics, halo conv)
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=150 TAIL=30 run compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cta_pairs_match and 1000"
T=150 TAIL=30 run compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "epilogue_configurations and 30000 and static"
T=150 TAIL=30 run compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_bn_backward_reduce_conv_dgrad and 30 and static"
