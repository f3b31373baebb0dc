#!/bin/bash
# A/B builds of the library (build_variants/lib*.so), one multi-frame launch per run, on the BASELINE workload families:
# noise 1080p RGB, jitter c=5 / c=10 4K gray, static 8K gray (young stacks: first 32 frames; aged stacks: frames 600..631).
# Usage (under gpurun): bash tools/gpu_ab.sh [quick]
set -u
for so in build_variants/lib*.so; do
  echo "#### $so"
  run() { ADDER_B200_SO=$PWD/$so timeout 300 python tools/profile_run.py --reps 3 --count --batch --ignore-errors "$@" 2>&1 | grep -E "counted|rep 2|rror" | sed -e 's/^/   /'; }
  echo " noise 1080p rgb";   run --frames 48 --cap 2
  echo " jitter 4k c=10";    run --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 48 --cap 2
  echo " static 8k young";   run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5
  [ "${1:-}" = "quick" ] && continue
  echo " jitter 4k c=5";     run --w 3840 --h 2160 --c 1 --kind 2 --manual 5 --frames 48 --cap 2
  echo " static 8k aged";    run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 608
done
