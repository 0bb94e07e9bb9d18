#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=700 TAIL=60 run python -m pytest tests -m gpu -q
T=300 TAIL=30 run python scripts/debug/dbg_bias_h2048.py
T=400 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3 --dump-gemm-profile gpurun_out/r02_gemm_launches_a.json
# launch list of one step (shares), then full captures of the three worst GEMM classes come in a later call
T=500 TAIL=3 run ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1700 --csv --log-file gpurun_out/r02_launches_a.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-incumbent
