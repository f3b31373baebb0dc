#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== ncu launch list of bench.py (bounded: 250 frames per step, no side workloads)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02q_bench_launches.csv python bench.py --frames 250 --steps 2 --warmup 3 --no-cpu --no-workloads --traffic off > gpurun_out/r02q_ncu_bench.log 2>&1; tail -c 200 gpurun_out/r02q_ncu_bench.log
echo "== ncu full: cfg5 aged stacks (the headline kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 37 -c 1 -f -o gpurun_out/r02q_static_prof python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 16 --cap 0.25 --warm-frames 592 --batch --reps 1 > gpurun_out/r02q_static_ncu.log 2>&1; tail -1 gpurun_out/r02q_static_ncu.log
echo "== ncu full: cfg2 noise"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 1 -f -o gpurun_out/r02q_prof python tools/profile_run.py --frames 16 --cap 2 --batch --reps 3 > gpurun_out/r02q_ncu.log 2>&1; tail -1 gpurun_out/r02q_ncu.log
echo "== ncu full: cfg3 c=10"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:integrate_frame -s 2 -c 1 -f -o gpurun_out/r02q_jit10_prof python tools/profile_run.py --w 3840 --h 2160 --c 1 --kind 2 --manual 10 --frames 16 --cap 2 --batch --reps 3 > gpurun_out/r02q_jit10_ncu.log 2>&1; tail -1 gpurun_out/r02q_jit10_ncu.log
echo "== compute-sanitizer on the exchange + compact kernels (small case)"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_multi.py tests/test_compact_form.py -m gpu -x -q -k "one_process or expands" 2>&1 | tail -6 > gpurun_out/r02q_sanitizer.txt; cat gpurun_out/r02q_sanitizer.txt
