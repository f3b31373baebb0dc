#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02l_ab_pf3.txt
echo "== parity pf3 multiframe/fullsize subset"
ADDER_B200_SO=$PWD/build_variants/lib_pf3.so timeout 900 python -m pytest tests/test_gpu_multiframe.py tests/test_gpu_fullsize.py -m gpu -x -q -k "multi or cfg5 or long_integration" 2>&1 | tail -3
