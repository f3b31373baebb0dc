#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02m_ab_fastwalk.txt
echo "== parity fast walk"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multiframe.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "64_frames or long_integration or sweep" 2>&1 | tail -3
