#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== whole GPU suite (pair-record state layout, exchange, compact form, C consumer)"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -16
echo "== A/B"
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02e_ab.txt
echo "== bench full"
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'dram/alg', d['roofline'].get('traffic_over_algorithmic'), 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('parse failed', e)
PY
tail -20 gpurun_out/r02e_bench.err
