#!/bin/bash
# Quick GPU visit: smoke, then the profiling driver (CUDA-event time per frame) under a few env variants.
# Usage (under gpurun): bash tools/gpu_quick.sh "VAR=val VAR2=val" "VAR=val" ...   (one quoted env set per variant; "" = defaults)
set -u
mkdir -p gpurun_out
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
[ "${PIPESTATUS[0]}" = "0" ] || { echo "smoke failed or hung: stopping"; exit 1; }
for v in "$@"; do
  echo "== variant [$v]"
  env $v timeout 300 python tools/profile_run.py --reps 3 ${PROFILE_ARGS:-} 2>&1 | tail -2
done
