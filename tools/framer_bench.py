#!/usr/bin/env python
"""The INSTANTANEOUS framer's ingest on its own: NF frames of 1080p gray noise are transcoded
into NF event buffers in HBM, then ingested back to back (adder_b200_framer_ingest_events_device_async, one wait at
the end).  Prints us per frame on the host clock (the framer has its own stream) and the share of the HBM roofline of
the records + pixel state it moves.  Under ncu: `-k regex:framer_ingest -s 4 -c 1`.
usage: python tools/framer_bench.py [--frames N] [--chunk-rows R]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adder_codec_rs_b200 as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--chunk-rows", type=int, default=1)
ap.add_argument("--kind", type=int, default=1)
a = ap.parse_args()
W, H, NF, REF, DTM, PEAK = 1920, 1080, a.frames, 255, 7650, 6540.2
P = W * H
v = A.Video(W, H, 1)
v.time_parameters(REF * 30, REF, DTM, None)
v.update_crf(3)
d_gray = v.device_alloc(P * NF)
v.synth_frames(d_gray, 0, NF, a.kind, 0xADDE5)
cap = P * 2
d_evs = [v.device_alloc(cap * 12) for _ in range(NF)]
d_ofs = [v.device_alloc((v.n_chunks + 1) * 4) for _ in range(NF)]
for f in range(NF):
    v.integrate_frames_device(d_gray.ptr + f * P, P, 1, float(REF), d_evs[f].ptr, cap, d_ofs[f].ptr)
v.sync()
n_events = sum(int(d_ofs[f].to_host(np.uint32)[-1]) for f in range(NF))
for rep in range(3):
    fr = A.Framer(W, H, 1, a.chunk_rows, 3, A.TIME_ABSOLUTE_T, REF * 30, REF, DTM, output_fps=30.0, ring_frames=160)
    t0 = time.perf_counter()
    for f in range(NF):
        fr.ingest_events_device_async(d_evs[f].ptr, d_ofs[f].ptr)
    fr.frame_ready()
    dt = time.perf_counter() - t0
    bytes_frame = (12 + 2 * 16) * n_events / NF
    print(f"rep {rep}: {dt / NF * 1e6:7.1f} us/frame, {n_events / NF:.0f} events/frame, {bytes_frame / (dt / NF) / 1e9:6.0f} GB/s of records + pixel state "
          f"= {bytes_frame / (dt / NF) / 1e9 / PEAK:.3f} of {PEAK:.0f}")
