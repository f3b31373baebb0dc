#!/usr/bin/env python
"""Small driver for ncu: the bench workload's plane (1920x1080x3 noise, crf 3) for a few dozen frames,
device-resident, so that `ncu -k regex:integrate_frame -s <skip> -c <n>` lands on steady-state frames.
Prints per-frame CUDA-event time for reference (never a bench number when run under a profiler)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adder_codec_rs_b200 as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--c", type=int, default=3)
ap.add_argument("--frames", type=int, default=48)
ap.add_argument("--kind", type=int, default=1)
ap.add_argument("--crf", type=int, default=3)
ap.add_argument("--ref", type=int, default=255)
ap.add_argument("--dtm", type=int, default=7650)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()

P = a.w * a.h * a.c
v = A.Video(a.w, a.h, a.c)
assert v.time_parameters(a.ref * 30, a.ref, a.dtm, None)
v.update_crf(a.crf)
d_frames = v.device_alloc(P * a.frames)
v.synth_frames(d_frames, 0, a.frames, a.kind, 0xADDE5)
stride = P * 2
d_events = v.device_alloc(stride * 12 * 4)
v.sync()
for rep in range(a.reps):
    v.reset_state()
    v.update_crf(a.crf)
    v.timer_start()
    for f in range(a.frames):
        v.integrate_frames_device(d_frames.ptr + f * P, P, 1, float(a.ref), d_events.ptr + (f % 4) * stride * 12, stride, None)
    ms = v.timer_stop()
    v.sync()
    print(f"rep {rep}: {a.frames} frames {ms:.3f} ms  -> {ms / a.frames * 1e3:.1f} us/frame, {P * a.frames / ms / 1e3:.1f} Mpx/s, events {v.events_emitted()}")
