#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=900 TAIL=12 run python -m pytest tests -m gpu -q -k "fused_bn or cta_pairs or backbone or trainer or full_size"
B="python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5"
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_BNR_PREFETCH=1 VTX_BNR_EARLY=0 $B
T=400 TAIL=1 run env VTX_BNR_PREFETCH=2 VTX_BNR_EARLY=0 $B
T=400 TAIL=1 run env VTX_BNR_PREFETCH=1 VTX_BNR_EARLY=1 $B
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_BNR_PREFETCH=1 VTX_BNR_EARLY=0 $B
T=400 TAIL=1 run env VTX_BNR_FUSE=0 $B
