#!/usr/bin/env python
"""adder_b200_expand_compact on the host's threads (no device involved): one 1080p RGB frame's dense compact block
(0.94 events per pixel-channel) and a sparse one (1/8) -> 12-byte records; ms per frame for 1, 4, 8, 16 threads."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adder_codec_rs_b200 import binding as B  # noqa: E402

W, H, C = 1920, 1080, 3
P = W * H * C
rng = np.random.default_rng(0)
counts = (rng.random(P) < 0.94).astype(np.uint8)
E = int(counts.sum())
block = np.concatenate([counts, rng.integers(0, 256, 5 * E, dtype=np.uint8)])
out = np.empty(E, dtype=B.EVENT_DTYPE)
idx = np.sort(rng.choice(P, P // 8, replace=False)).astype(np.uint32)
Es = len(idx)
blk = np.zeros(9 * Es, dtype=np.uint8)
v = blk.reshape(Es, 9)
v[:, 0:4] = idx.view(np.uint8).reshape(Es, 4)
v[:, 4:] = rng.integers(0, 256, (Es, 5), dtype=np.uint8)
outs = np.empty(Es, dtype=B.EVENT_DTYPE)
print(f"# {len(os.sched_getaffinity(0))} host cores")
for name, b, n, o in (("dense", block, E, out), ("sparse", blk, Es, outs)):
    for T in (1, 4, 8, 16):
        B.expand_compact(W, H, C, 0, b, n, o, T)
        t0 = time.perf_counter()
        for _ in range(5):
            B.expand_compact(W, H, C, 0, b, n, o, T)
        dt = (time.perf_counter() - t0) / 5
        print(f"{name:6} {T:2d} threads: {dt * 1e3:6.2f} ms per frame, {P / dt / 1e6:6.0f} Mpx/s, {n / dt / 1e6:5.0f} Mevents/s")
