#!/bin/bash
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=400 TAIL=30 run python -m pytest tests/test_gpu_kernels.py -m gpu -q
T=500 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5 --dump-gemm-profile gpurun_out/r02l_gemm_launches_dual_all.json
T=900 TAIL=30 run python -m pytest tests/test_gpu_parity.py -m gpu -q
sed -i 's/#define VTX_DUAL_EPI 2/#define VTX_DUAL_EPI 1/' virtex_b200/csrc/gemm_tc.cu
T=600 TAIL=2 run python -m virtex_b200.build
T=500 TAIL=1 run python bench.py --skip-cpu --skip-incumbent --steps 30 --warmup 5 --dump-gemm-profile gpurun_out/r02l_gemm_launches_dual_narrow.json
