#!/bin/bash
# Dynamic tile scheduler: GPU test suite, then bench A/B (dynamic vs VTX_GEMM_STATIC=1) and an ncu look at the BN / attention kernels
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=1500 TAIL=15 run python -m pytest tests -m gpu -x -q
T=400 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5 --dump-gemm-profile gpurun_out/r02m_gemm_launches_dynamic.json
T=400 TAIL=1 run env VTX_GEMM_STATIC=1 python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5 --dump-gemm-profile gpurun_out/r02m_gemm_launches_static.json
T=400 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5
T=400 TAIL=1 run env VTX_GEMM_STATIC=1 python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5
T=600 TAIL=3 run ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --clock-control none -k regex:"bn_bwd|bn_act|attn_|maxpool" -o gpurun_out/r02m_bn_kernels python scripts/profile_step.py
