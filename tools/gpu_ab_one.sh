#!/bin/bash
# One library (ADDER_B200_SO or the in-tree one) on the BASELINE workload families; extra env passes through.
# Usage (under gpurun): [ADDER_B200_SO=...] bash tools/gpu_ab_one.sh [quick]
set -u
run() { timeout 90 python tools/profile_run.py --reps 3 --count --batch --ignore-errors "$@" 2>&1 | grep -E "counted|rep 2|rror" | sed -e 's/^/   /'; }
echo " noise 1080p rgb";   run --frames 48 --cap 2
echo " jitter 4k c=10";    run --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 48 --cap 2
echo " static 8k aged";    run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 608
[ "${1:-}" = "quick" ] && exit 0
echo " jitter 4k c=0";     run --w 3840 --h 2160 --c 1 --kind 2 --manual 0 --frames 48 --cap 2
echo " jitter 4k c=5";     run --w 3840 --h 2160 --c 1 --kind 2 --manual 5 --frames 48 --cap 2
echo " static 8k young";   run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5
