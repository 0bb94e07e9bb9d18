#!/bin/bash
# Round-2 evidence batch at one commit: tests, smoke, bench (all three BASELINE model configs, with the eager incumbent),
# reference arm, launch list, GEMM per-launch profile and full ncu captures.  TAG names the output files.
set -u
TAG=${TAG:-r02s}
mkdir -p gpurun_out
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }
T=1200 TAIL=15 run python -m pytest tests -m gpu -q
T=300 TAIL=5 run python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
T=900 TAIL=1 run python bench.py --steps 20 --warmup 5 --dump-gemm-profile gpurun_out/${TAG}_gemm_launches.json
T=600 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3 --config depth_ablations/bicaptioning_R_50_L4_H1024.yaml
T=700 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3 --config backbone_ablations/bicaptioning_R_101_L1_H1024.yaml --config-override MODEL.TEXTUAL.NAME transdec_postnorm::L1_H2048_A32_F8192
T=600 TAIL=1 run python bench.py --impl reference --steps 2 --warmup 1
T=600 TAIL=3 run ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_step.csv python scripts/profile_step.py
T=600 TAIL=3 run ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 7 -c 7 -o gpurun_out/${TAG}_gemm_cases python scripts/ncu_gemm_cases.py l1conv3 l1dgrad_bnr l1conv3_dgrad_bnr l1conv l3conv vocab_pair ffn2_pair
