#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== the two new host-form tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resumes or normal_mode_long" 2>&1 | tail -5
echo "== A/B"
bash tools/gpu_ab.sh quick 2>&1 | tee gpurun_out/r02b_ab.txt
