#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity with R = 8 forced (the plain instantiation serves every default-mode case)"
ADDER_B200_R=8 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multiframe.py -m gpu -x -q 2>&1 | tail -3
echo "== whole GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== A/B plain on / off (same library)"
for plain in 0 1; do
  echo "#### ADDER_B200_PLAIN=$plain"
  run() { ADDER_B200_PLAIN=$plain timeout 300 python tools/profile_run.py --reps 3 --count --batch "$@" 2>&1 | grep -E "rep 2|rror" | sed -e 's/^/   /'; }
  echo " noise 1080p rgb";   run --frames 48 --cap 2
  echo " jitter 4k c=10";    run --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 48 --cap 2
  echo " jitter 4k c=5";     run --w 3840 --h 2160 --c 1 --kind 2 --manual 5 --frames 48 --cap 2
  echo " static 8k aged";    run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 608
done 2>&1 | tee gpurun_out/r02o_ab_plain.txt
