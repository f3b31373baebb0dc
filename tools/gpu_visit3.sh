#!/bin/bash
# ncu source-level captures of the lean-only experiment kernel on noise and static-aged
set -u
mkdir -p gpurun_out
export ADDER_B200_SO=$PWD/build_variants/lib_leanonly.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 1 -c 1 -f -o gpurun_out/r02c_leanonly_noise python tools/profile_run.py --frames 16 --cap 2 --batch --reps 2 --ignore-errors > gpurun_out/r02c_noise.log 2>&1; tail -2 gpurun_out/r02c_noise.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 13 -c 1 -f -o gpurun_out/r02c_leanonly_static python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 16 --cap 0.5 --warm-frames 192 --batch --reps 1 --ignore-errors > gpurun_out/r02c_static.log 2>&1; tail -2 gpurun_out/r02c_static.log
ls -la gpurun_out/r02c*
