#!/bin/bash
# Round-2 helper: validate the opt-in experimental features on a B200, one at a time (run under gpurun).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/validate_experimental.sh > gpurun_out/validate_x.log 2>&1; tail -40 gpurun_out/validate_x.log'
# Every step runs under its own `timeout` so that a hanging kernel cannot eat the box; a feature is good when its
# kernel tests AND the full parity suite pass with it enabled, and the bench line (config.experimental names it) is
# faster than the baseline line printed first.
set -u
cd "$(dirname "$0")/.."
run() { echo "=== $*"; timeout -k 5 "${T:-420}" "$@" 2>&1 | tail -${TAIL:-6}; echo "--- exit ${PIPESTATUS[0]}"; }

T=300 run python bench.py --skip-cpu --steps 10 --warmup 3
VTX_RUN_UNVERIFIED=1 run python -m pytest tests -m gpu -q -x -k "baseline_config or checkpoint_resume or edge_batch"
for f in head_x gemm_x backbone_x stem_s2d pdl; do
  export VTX_EXPERIMENTAL=$f
  T=200 run python -m pytest tests -m gpu -q -x -k "experimental"
  run python -m pytest tests -m gpu -q -x
  T=300 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3
  unset VTX_EXPERIMENTAL
done
export VTX_EXPERIMENTAL=all
T=300 TAIL=1 run python bench.py --skip-cpu --steps 10 --warmup 3
