#!/bin/bash
# usage: scripts/gpu_submit.sh <timeout_s> <script relative to repo root> [gpus]
# Freezes a copy of the working tree under .stage/ (git-ignored, shipped by gpurun) and runs the script from there, so
# edits made while the call waits in the pod's queue (every retry re-snapshots /root/repo) cannot leak into the run.
set -u
cd "$(dirname "$0")/.."
T=$1; S=$2; G=${3:-1}
rm -rf .stage && mkdir .stage
tar -c --exclude=./.git --exclude=./.stage --exclude=./gpurun_out --exclude=./build --exclude=./.pytest_cache --exclude=__pycache__ . | tar -x -C .stage
[ -f ".stage/$S" ] || { echo "staging failed: .stage/$S missing"; exit 1; }
CMD="cd .stage && mkdir -p gpurun_out && bash $S > gpurun_out/$(basename $S .sh).log 2>&1; mkdir -p ../gpurun_out && cp -r gpurun_out/. ../gpurun_out/; tail -3 gpurun_out/$(basename $S .sh).log"
GF=""; [ "$G" != "1" ] && GF="--gpus $G"
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $GF --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
