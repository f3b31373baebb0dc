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
ap.add_argument("--batch", action="store_true", help="hand all frames to one integrate_frames_device call (multi-frame launches)")
ap.add_argument("--cap", type=float, default=4, help="event records per pixel-channel per frame the output buffer holds")
ap.add_argument("--warm-frames", type=int, default=0, help="frames integrated (untimed) before the timed ones, so that the timed frames see aged pixel stacks; "
                "the timed frames are then frames warm..warm+frames of the sequence and the state is NOT reset between reps")
ap.add_argument("--manual", type=int, default=-1, help="quality_manual(c, c, dtm/ref, 1) instead of crf (BASELINE cfg 3 sweep)")
ap.add_argument("--normal", action="store_true", help="PixelMultiMode::Normal")
ap.add_argument("--offsets", action="store_true", help="also ask for the per-frame chunk offsets (as bench.py does)")
ap.add_argument("--ignore-errors", action="store_true", help="experiments with partial kernels: report the time even when sync() reports a device error")
ap.add_argument("--count", action="store_true", help="one extra untimed pass with the counting twin: algorithmic bytes + roofline fraction")
a = ap.parse_args()

P = a.w * a.h * a.c
v = A.Video(a.w, a.h, a.c)


def sync():
    try:
        v.sync()
    except A.AdderError as e:
        if not a.ignore_errors:
            raise
        print("  (device error ignored:", e, ")")


assert v.time_parameters(a.ref * 30, a.ref, a.dtm, None)
def quality():
    if a.manual >= 0:
        v.update_quality_manual(a.manual, a.manual, a.dtm // a.ref, 1, 0.0)
    else:
        v.update_crf(a.crf)


quality()
if a.normal:
    v.write_out(None, A.MULTI_NORMAL)
d_frames = v.device_alloc(P * a.frames)
stride = int(P * a.cap)
if a.warm_frames:
    assert a.batch
    d_tmp_ev = v.device_alloc(stride * 12 * a.frames)
    for f0 in range(0, a.warm_frames, a.frames):
        n = min(a.frames, a.warm_frames - f0)
        v.synth_frames(d_frames, f0, n, a.kind, 0xADDE5)
        v.integrate_frames_device(d_frames.ptr, P, n, float(a.ref), d_tmp_ev.ptr, stride, None)
    sync()
    d_tmp_ev.free()
v.synth_frames(d_frames, a.warm_frames, a.frames, a.kind, 0xADDE5)
d_events = v.device_alloc(stride * 12 * (a.frames if a.batch else 4))
d_off = v.device_alloc((v.n_chunks + 1) * 4 * a.frames) if a.offsets else None
sync()
alg = None
if a.count:
    v.set_counting(True)
    if a.batch:
        v.integrate_frames_device(d_frames.ptr, P, a.frames, float(a.ref), d_events.ptr, stride, None)
    else:
        for f in range(a.frames):
            v.integrate_frames_device(d_frames.ptr + f * P, P, 1, float(a.ref), d_events.ptr + (f % 4) * stride * 12, stride, None)
    sync()
    c = v.read_counters()
    v.set_counting(False)
    alg = (1 + 8 + 8) * P * a.frames + 16 * (c["node_loads"] + c["node_stores"]) + c["display_writes"] + 12 * c["events"]
    Lm = (c["live_nodes_in"] + c["live_nodes_out"]) / (2.0 * P * a.frames)
    print(f"SURVEY 8(d) formula: L {Lm:.3f} -> {1 + 2 * (12 + 16 * Lm) + 1 + 12 * c['events'] / (P * a.frames):.2f} B/px-frame")
    print(f"counted: {alg / (P * a.frames):.2f} B/px-frame, loads {c['node_loads'] / (P * a.frames):.3f} stores {c['node_stores'] / (P * a.frames):.3f} "
          f"display {c['display_writes'] / (P * a.frames):.3f} events {c['events'] / (P * a.frames):.3f} per px-frame")
for rep in range(a.reps):
    if not a.warm_frames:
        v.reset_state()
        quality()
    v.timer_start()
    if a.batch:
        v.integrate_frames_device(d_frames.ptr, P, a.frames, float(a.ref), d_events.ptr, stride, d_off.ptr if d_off else None)
    else:
        for f in range(a.frames):
            v.integrate_frames_device(d_frames.ptr + f * P, P, 1, float(a.ref), d_events.ptr + (f % 4) * stride * 12, stride, d_off.ptr if d_off else None)
    ms = v.timer_stop()
    sync()
    extra = f", {alg / ms / 1e6:.0f} GB/s = {alg / ms / 1e6 / 6540.2:.3f} of 6540" if alg else ""
    print(f"rep {rep}: {a.frames} frames {ms:.3f} ms  -> {ms / a.frames * 1e3:.1f} us/frame, {P * a.frames / ms / 1e3:.1f} Mpx/s, events {v.events_emitted()}{extra}")
