#!/usr/bin/env python
"""Time of the one exchange step of row-band sharding (sharding.gather_events over NCCL): every rank holds a 1080-row
band of RGB noise (the bench plane per GPU) and its frame's events in HBM; rank 0 receives the whole frame's stream in
raster order.  Launch: python -m torch.distributed.run --nproc-per-node N tools/gather_bench.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import adder_codec_rs_b200 as A  # noqa: E402
from adder_codec_rs_b200 import sharding as S  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, C, NF, REF = 1920, 1080, 3, 12, 255
v = S.BandedVideo(W, H * world, C, rank, world, device=local)
v.time_parameters(REF * 30, REF, 7650, None)
v.update_crf(3)
P = W * v.rows * C
d_frames = v.device_alloc(P * NF)
v.synth_frames(d_frames, 0, NF, 1, 0xADDE5 + rank)
cap = P * 2
d_events = v.device_alloc(cap * 12)
d_off = v.device_alloc((v.n_chunks + 1) * 4)
ts, kernel_ms = [], []
for f in range(NF):
    v.timer_start()
    v.integrate_frames_device(d_frames.ptr + f * P, P, 1, float(REF), d_events.ptr, cap, d_off.ptr)
    kernel_ms.append(v.timer_stop())
    v.sync()
    off = S.device_bytes_as_tensor(d_off.ptr, (v.n_chunks + 1) * 4, local).view(torch.int32).to(torch.int64)
    n = int(off[-1])
    ev = S.device_bytes_as_tensor(d_events.ptr, n * 12, local)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    g_ev, g_cc = S.gather_events(ev, off[1:] - off[:-1], dst=0)
    torch.cuda.synchronize()
    dist.barrier()
    ts.append(time.perf_counter() - t0)
    if rank == 0 and f == NF - 1:
        tot = g_ev.numel()
        print(f"N={world}: frame of {W}x{H * world}x{C}: {tot // 12} events ({tot / 1e6:.0f} MB) gathered on rank 0 in raster order: "
              f"{np.median(ts[2:]) * 1e3:.2f} ms median per frame (host clock, barrier to barrier; band kernel {np.median(kernel_ms[2:]) * 1e3:.0f} us) "
              f"-> {tot / np.median(ts[2:]) / 1e9:.0f} GB/s into rank 0")
dist.destroy_process_group()
