#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== A/B default L2 fetch granularity"
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02f_ab.txt
echo "== A/B ADDER_B200_L2_FETCH=32"
ADDER_B200_L2_FETCH=32 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02f_ab_l2_32.txt
echo "== dram traffic, aged static, granularity default / 32 (pair layout)"
for g in 64 32; do
ADDER_B200_L2_FETCH=$g ADDER_B200_SO=$PWD/build_variants/lib_pair.so timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:integrate_frame python tools/profile_run.py --batch --count --reps 1 --frames 16 --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --cap 0.25 --warm-frames 592 2>&1 | grep -E "counted|dram__" | tail -3
done
