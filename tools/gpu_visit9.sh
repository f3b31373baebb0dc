#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== framer / compact / parity subset"
timeout 900 python -m pytest tests/test_gpu_framer.py tests/test_compact_form.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "== next rows (framer timing)"
timeout 600 python tools/next_rows_bench.py > gpurun_out/r02h_next_rows.txt 2>&1; grep -E "row 3|CPU port, ingest" gpurun_out/r02h_next_rows.txt
echo "== bench smoke with the cfg2 e2e legs"
timeout 900 python bench.py --frames 64 --steps 2 --warmup 3 --cpu-seconds 3 --traffic off > gpurun_out/r02h_bench_smoke.json 2> gpurun_out/r02h_bench_smoke.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02h_bench_smoke.json').read().strip().splitlines() if l.startswith('{')][-1])
    for e in d['workloads']:
        if 'e2e' in e: print(json.dumps(e['e2e'])[:1500])
except Exception as e: print('parse failed', e)
PY
tail -3 gpurun_out/r02h_bench_smoke.err
echo "== ncu launch list of bench.py (bounded: 250 frames per step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02h_bench_launches.csv python bench.py --frames 250 --steps 2 --warmup 3 --no-cpu --no-workloads --traffic off > gpurun_out/r02h_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r02h_ncu_bench.log
echo "== ncu full: cfg5 aged stacks"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 37 -c 1 -f -o gpurun_out/r02h_static_prof python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 16 --cap 0.25 --warm-frames 592 --batch --reps 1 > gpurun_out/r02h_static_ncu.log 2>&1; tail -2 gpurun_out/r02h_static_ncu.log
echo "== ncu full: cfg3 c=10"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 1 -f -o gpurun_out/r02h_jit10_prof python tools/profile_run.py --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 16 --cap 2 --batch --reps 3 > gpurun_out/r02h_jit10_ncu.log 2>&1; tail -2 gpurun_out/r02h_jit10_ncu.log
echo "== ncu full: cfg2 noise"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 1 -f -o gpurun_out/r02h_prof python tools/profile_run.py --frames 16 --cap 2 --batch --reps 3 > gpurun_out/r02h_ncu.log 2>&1; tail -2 gpurun_out/r02h_ncu.log
ls -la gpurun_out/r02h*
