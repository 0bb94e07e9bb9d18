#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; rc=${PIPESTATUS[0]}; echo "--- exit $rc"; return $rc; }
T=300 TAIL=120 run python scripts/debug/diag_engine_bnr.py 6
T=300 TAIL=120 run python scripts/debug/diag_engine_bnr.py 64
