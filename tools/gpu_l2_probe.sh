#!/bin/bash
# Does the integrate kernel run faster per pixel when the whole plane's state stays in L2?  Aged static stacks (dtm 2^20, 608 warm
# frames) on planes from 0.5 to 33 Mpx at a fixed tile size (ADDER_B200_R): the upper bound of what keeping a tile's state in L2
# across frames (T-frame blocking) could buy.  Usage (under gpurun): bash tools/gpu_l2_probe.sh
set -u
for r in 4 8; do
for wh in "1024 512" "1024 1024" "2048 1024" "4096 2048" "7680 4320"; do
  set -- $wh
  echo "== R=$r plane $1 x $2"
  ADDER_B200_R=$r timeout 200 python tools/profile_run.py --reps 3 --count --batch --ignore-errors --w $1 --h $2 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 608 2>&1 | grep -E "counted|rep 2|rror" | sed -e 's/^/   /'
done
done
