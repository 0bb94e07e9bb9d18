#!/bin/bash
# bn3 fusion threshold A/B: layers 1-2 only (default) vs every identity-followed block
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
B="python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5"
T=300 TAIL=1 run $B
T=300 TAIL=1 run env VTX_BNR_BN3_MIN_ROWS=0 $B
T=300 TAIL=1 run $B
T=300 TAIL=1 run env VTX_BNR_BN3_MIN_ROWS=0 $B
