#!/bin/bash
# A/B builds of the library (build_variants/lib*.so), one multi-frame launch per run, on three workloads:
# noise 1080p RGB, jitter c=10 4K gray, static 8K gray.
set -u
for so in build_variants/lib*.so; do
  echo "#### $so"
  ADDER_B200_SO=$PWD/$so timeout 300 python tools/profile_run.py --reps 3 --frames 48 --cap 2 --batch 2>&1 | tail -1
  ADDER_B200_SO=$PWD/$so timeout 300 python tools/profile_run.py --reps 3 --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 48 --cap 2 --batch 2>&1 | tail -1
  ADDER_B200_SO=$PWD/$so timeout 300 python tools/profile_run.py --reps 3 --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 1 --batch 2>&1 | tail -1
done
