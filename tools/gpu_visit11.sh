#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity of the L1 variants on multi-frame launches"
for so in l1_pf2; do ADDER_B200_SO=$PWD/build_variants/lib_$so.so timeout 900 python -m pytest tests/test_gpu_multiframe.py tests/test_gpu_fullsize.py -m gpu -x -q -k "multi or 64_frames or long" 2>&1 | tail -3; done
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02j_ab_l1.txt
