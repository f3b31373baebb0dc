#!/bin/bash
# ncu --set full of one 16-frame launch on aged 8K stacks and on 1080p noise.  Usage (under gpurun): [ADDER_B200_SO=..] bash tools/gpu_ncu_two.sh <tag> [static|noise|jit10 ...]
set -u
TAG=$1; shift
WHAT=${*:-static noise}
mkdir -p gpurun_out
for w in $WHAT; do
  case $w in
    static) timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 37 -c 1 -f -o gpurun_out/${TAG}_static_prof python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 16 --cap 0.25 --warm-frames 592 --batch --reps 1 --ignore-errors > gpurun_out/${TAG}_static_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_static_ncu.log;;
    noise) timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 1 -f -o gpurun_out/${TAG}_prof python tools/profile_run.py --frames 16 --cap 2 --batch --reps 3 --ignore-errors > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log;;
    jit10) timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 1 -f -o gpurun_out/${TAG}_jit10_prof python tools/profile_run.py --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 16 --cap 2 --batch --reps 3 --ignore-errors > gpurun_out/${TAG}_jit10_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_jit10_ncu.log;;
  esac
done
ls -la gpurun_out | grep ${TAG}
