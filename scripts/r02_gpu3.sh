#!/bin/bash
# Round-2 measurement batch: tests, bench (all three BASELINE model configs), launch list and full ncu captures.
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=900 TAIL=60 run python -m pytest tests -m gpu -q
T=300 TAIL=30 run python scripts/debug/dbg_bias_h2048.py
T=500 TAIL=1 run python bench.py --steps 20 --warmup 5 --dump-gemm-profile gpurun_out/r02_gemm_launches.json
cp gpurun_out/r02_gemm_launches.json gpurun_out/r02_gemm_launches_c2.json 2>/dev/null
T=500 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3 --config depth_ablations/bicaptioning_R_50_L4_H1024.yaml --dump-gemm-profile gpurun_out/r02_gemm_launches_c4.json
T=600 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3 --config backbone_ablations/bicaptioning_R_101_L1_H1024.yaml --config-override MODEL.TEXTUAL.NAME transdec_postnorm::L1_H2048_A32_F8192 --dump-gemm-profile gpurun_out/r02_gemm_launches_c5.json
# one optimisation step, every launch with its device time and DRAM bytes (cold-cache, serialised: compare shares)
T=600 TAIL=3 run ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py
# full captures: the worst GEMM classes and the BN passes
T=600 TAIL=3 run ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 5 -o gpurun_out/r02_gemm_cases python scripts/ncu_gemm_cases.py l1conv3 l1dgrad l1conv l2wgrad l3conv
T=600 TAIL=3 run ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bn_bwd_apply|bn_bwd_reduce|bn_act|attn_" -c 12 -o gpurun_out/r02_bn_attn python scripts/profile_step.py
