#!/bin/bash
# round-2 visit 1: the new full-size tests on the round-1 kernel, then the whole GPU suite on the current kernel, then A/B
set -u
mkdir -p gpurun_out
echo "== new tests on the round-1 kernel"
ADDER_B200_SO=$PWD/build_variants/lib_r01.so timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --durations=8 2>&1 | tail -14
echo "== whole GPU suite, current kernel"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== A/B"
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02a_ab.txt
