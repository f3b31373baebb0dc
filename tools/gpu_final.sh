#!/bin/bash
# final visit of the round: smoke, the whole GPU suite, the bench exactly as the driver runs it, the reference arm, the ncu launch list of bench.py
set -u
TAG=${1:-r02v}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt
echo "== smoke"; timeout -k 10 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
echo "== bench (driver arguments)"
t0=$(date +%s); timeout 2400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$? seconds=$(( $(date +%s) - t0 ))"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',round(r['frac'],3),'dram/alg',r.get('traffic_over_algorithmic'),'cpu',round(d['cpu_baseline']['value']),'cores',d['cpu_baseline']['cores'],'launches',d['gpu_launches'],'clocks',d['clocks'])
PY
cat gpurun_out/${TAG}_bench.err | tail -18
echo "== reference arm"
t0=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_reference.json; echo "rc=$? seconds=$(( $(date +%s) - t0 ))"; cut -c1-200 gpurun_out/${TAG}_reference.json
echo "== ncu launch list of bench.py (device-resident steps first; -c bounds the e2e legs' launches)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-workloads --traffic off > gpurun_out/${TAG}_ncu_bench.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_bench.log
