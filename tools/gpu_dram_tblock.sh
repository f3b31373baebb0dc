#!/bin/bash
# DRAM bytes of one 32-frame launch on aged 8K stacks for T = 1 / 8 / 16 (offset form), and eager T = 1.
for cfg in "1 1" "8 1" "16 1" "1 0"; do
  set -- $cfg
  echo "TBLOCK=$1 OFFSET=$2"
  ADDER_B200_TBLOCK=$1 ADDER_B200_OFFSET=$2 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:integrate_frame -s 19 -c 1 python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.25 --warm-frames 608 --batch --reps 1 2>&1 | grep -E "dram__|gpu__time|lts__" | sed -e 's/^/   /'
done
