#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity with the long-integration variant forced on every case (R = 8, deep walk)"
ADDER_B200_R=8 ADDER_B200_DEEP=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multiframe.py -m gpu -x -q 2>&1 | tail -5
echo "== full size cfg5 + long integration"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "cfg5 or long_integration" 2>&1 | tail -5
echo "== A/B deep walk off / on"
for deep in 0 1; do
  echo "#### ADDER_B200_DEEP=$deep"
  run() { ADDER_B200_DEEP=$deep timeout 300 python tools/profile_run.py --reps 3 --count --batch "$@" 2>&1 | grep -E "counted|rep 2|rror" | sed -e 's/^/   /'; }
  echo " static 8k young";   run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5
  echo " static 8k aged";    run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 608
  echo " static 8k old";     run --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 1792
  echo " jitter 4k c=10";    run --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 48 --cap 2
  echo " noise 1080p rgb";   run --frames 48 --cap 2
done 2>&1 | tee gpurun_out/r02i_ab_deep.txt
