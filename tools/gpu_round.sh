#!/bin/bash
# One GPU visit: parity tests, a short bench, the ncu launch list and one full capture of the hot kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt
echo "== smoke"; timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "smoke failed or hung: stopping"; exit 1; }
[ "${PIPESTATUS[0]}" = "0" ] || { echo "smoke failed or hung: stopping"; exit 1; }
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== profile_run (plain: one launch per frame, then one launch for all frames)"; timeout 300 python tools/profile_run.py --reps 3 --cap 2 2>&1 | tail -3; timeout 300 python tools/profile_run.py --reps 3 --cap 2 --batch 2>&1 | tail -3
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py --frames 16 --cap 2 --batch --reps 4 > gpurun_out/${TAG}_ncu_list.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_list.log
echo "== ncu launch list of bench.py itself (the device-resident steps come first; -c bounds the e2e legs' per-frame launches)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_bench.log
echo "== ncu full"
# the integrate kernel as bench.py launches it: one launch spanning many frames (16 here, so that ncu can save and restore the device memory between replays)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 2 -f -o gpurun_out/${TAG}_prof python tools/profile_run.py --frames 16 --cap 2 --batch --reps 4 > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -12
