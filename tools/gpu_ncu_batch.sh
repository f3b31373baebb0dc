#!/bin/bash
# ncu --set full captures of one multi-frame launch (16 frames) on the two workloads the kernel is weakest on:
# 4K gray jitter c=10 and 8K gray static + blips.  Usage (under gpurun): bash tools/gpu_ncu_batch.sh <tag>
set -u
TAG=$1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 1 -c 1 -f -o gpurun_out/${TAG}_jit10_prof python tools/profile_run.py --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 16 --cap 2 --batch --reps 2 > gpurun_out/${TAG}_jit10_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_jit10_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 1 -c 1 -f -o gpurun_out/${TAG}_static8k_prof python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 16 --cap 1 --batch --reps 2 > gpurun_out/${TAG}_static8k_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_static8k_ncu_full.log
ls -la gpurun_out | tail -6
