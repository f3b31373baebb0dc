#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== A/B deep-level prefetch on / off"
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02g_ab.txt
echo "== dram traffic, aged static"
for so in pair pair_pf0; do for g in 64 32; do
echo "-- $so L2_FETCH=$g"
ADDER_B200_L2_FETCH=$g ADDER_B200_SO=$PWD/build_variants/lib_$so.so timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:integrate_frame python tools/profile_run.py --batch --count --reps 1 --frames 16 --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --cap 0.25 --warm-frames 592 2>&1 | grep -E "dram__" | tail -2 | awk -F'","' '{print $13, $NF}'
done; done
