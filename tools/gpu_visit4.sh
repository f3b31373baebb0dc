#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== exchange tests (one process)"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6
echo "== bench smoke (64 frames)"
timeout 900 python bench.py --frames 64 --steps 2 --warmup 3 --cpu-seconds 3 > gpurun_out/r02d_bench_smoke.json 2> gpurun_out/r02d_bench_smoke.err; echo "rc=$?"; tail -c 1500 gpurun_out/r02d_bench_smoke.json; tail -20 gpurun_out/r02d_bench_smoke.err
echo "== bench full"
/usr/bin/time -v timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic_over_algorithmic'), d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('parse failed', e)
PY
grep -E "Elapsed|Maximum resident" gpurun_out/r02d_bench.err; grep -v "^\s" gpurun_out/r02d_bench.err | tail -25
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 | cut -c1-400
