#!/bin/bash
# diagnostics of the masked-residual epilogue, CTA-pair (cta_group::2) tests, full suite, bench A/B pairs on / off
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=200 TAIL=60 run python scripts/debug/diag_bnr.py
if T=300 TAIL=25 run python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "cta_pairs"; then
  echo "PAIRS OK"
else
  echo "PAIRS FAILED: continuing with VTX_GEMM_PAIR=0"; nvidia-smi --query-gpu=name,memory.used --format=csv
  export VTX_GEMM_PAIR=0
fi
T=1500 TAIL=15 run python -m pytest tests -m gpu -x -q
B="python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5"
T=400 TAIL=1 run $B --dump-gemm-profile gpurun_out/r02o_gemm_launches_pair.json
T=400 TAIL=1 run env VTX_GEMM_PAIR=0 $B --dump-gemm-profile gpurun_out/r02o_gemm_launches_nopair.json
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_GEMM_PAIR=0 $B
