#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=1500 TAIL=30 run python -m pytest tests -m gpu -q
B="python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5"
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_PDL_AUX=0 $B
T=400 TAIL=1 run $B
T=400 TAIL=1 run env VTX_PDL_AUX=0 $B
