#!/bin/bash
# One ncu --set full capture of the integrate kernel on a chosen workload.
# Usage (under gpurun): bash tools/gpu_ncu.sh <tag> <profile_run.py args...>
set -u
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 30 -c 2 -f -o gpurun_out/${TAG}_prof python tools/profile_run.py "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
