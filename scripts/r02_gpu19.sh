#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=300 TAIL=80 run python scripts/debug/diag_bnr.py
T=300 TAIL=25 run python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cta_pairs_implicit"
export VTX_BNR_FUSE=0
T=1500 TAIL=15 run python -m pytest tests -m gpu -q -k "not fused_bn and not cta_pairs_match"
B="python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5"
T=400 TAIL=1 run $B --dump-gemm-profile gpurun_out/r02o_gemm_launches_pair.json
T=400 TAIL=1 run env VTX_GEMM_PAIR=0 $B --dump-gemm-profile gpurun_out/r02o_gemm_launches_nopair.json
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_GEMM_PAIR=0 $B
