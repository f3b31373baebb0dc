#!/bin/bash
# multi-GPU visit: usage (under gpurun --gpus N): bash tools/gpu_visit_multi.sh N [frames] [steps]
set -u
N=${1:-2}; FR=${2:-200}; ST=${3:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== exchange between processes (CUDA IPC over NVLink), sharding tests"
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -v 2>&1 | grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed" | sed -e 's/^tests.test_gpu_multi.py:://' > gpurun_out/r02g_multi_tests_n$N.txt; tail -4 gpurun_out/r02g_multi_tests_n$N.txt
echo "== bench --gpus $N --frames $FR (strong scaling, gather legs)"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --frames $FR --steps $ST --warmup 3 > gpurun_out/r02g_bench_n$N.json 2> gpurun_out/r02g_bench_n$N.err; echo "rc=$?"
grep "NCCL INFO" gpurun_out/r02g_bench_n$N.json | grep -E "NCCL version|nranks|NVLS multicast|via P2P" | head -40 > gpurun_out/r02g_nccl_n$N.txt; wc -l gpurun_out/r02g_nccl_n$N.txt
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02g_bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
    print('gather', d.get('gather',{}).get('value'), 'cfg4', d.get('gather_cfg4'))
except Exception as e: print('parse failed', e)
PY
grep -v "NCCL INFO" gpurun_out/r02g_bench_n$N.err | tail -15
echo "== reference arm under torchrun (cores must be > 1)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | cut -c1-250
