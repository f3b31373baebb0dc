#!/bin/bash
# The BASELINE config families through the profiling driver (device-resident, one launch for all frames, CUDA events,
# counted bytes).
# Usage: bash tools/gpu_workloads.sh [short]   (env vars such as ADDER_B200_DEEP_PREFETCH pass through)
set -u
run() { echo "== $*"; timeout 600 python tools/profile_run.py --reps 2 --count --batch --cap 2 "$@" 2>&1 | tail -4; }
run --w 1920 --h 1080 --c 3 --kind 1 --crf 3 --frames 48
run --w 3840 --h 2160 --c 1 --kind 2 --manual 0 --frames 48
run --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 48
run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32
[ "${1:-}" = "short" ] && exit 0
run --w 3840 --h 2160 --c 1 --kind 2 --manual 5 --frames 48
run --w 3840 --h 2160 --c 3 --kind 1 --crf 3 --frames 24
run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --normal
run --w 640 --h 480 --c 1 --kind 0 --dtm 255 --frames 30 --cap 4
