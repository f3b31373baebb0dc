#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (hoisted fire) on the cases + multiframe + fullsize subset"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multiframe.py tests/test_gpu_framer.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "64_frames or long_integration or sweep" 2>&1 | tail -3
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02k_ab_hoist.txt
echo "== next rows"
timeout 600 python tools/next_rows_bench.py > gpurun_out/r02k_next_rows.txt 2>&1; grep -E "row 3|asynchronous|CPU port, ingest" gpurun_out/r02k_next_rows.txt
