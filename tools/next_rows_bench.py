#!/usr/bin/env python
"""Timing of the kernels either side of the integrate kernel at 1080p, next to the oracle's
CPU port of the same step on this box's host cores.  Device times are CUDA events on the handle's stream unless the
call itself synchronises (the framer returns is_frame_0_filled, so its calls are timed on the host clock, sync
included).  Not a bench line: supporting numbers for DESIGN.md §4.2 (never run under a profiler when quoted)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adder_codec_rs_b200 as A  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

W, H, NF, REF, DTM, PEAK = 1920, 1080, 24, 255, 7650, 6540.2
quick = "--quick" in sys.argv  # device side only, fewer frames (for the ncu launch list)
if quick:
    NF = 6


def video(c, src_c=None):
    v = A.Video(W, H, c)
    v.time_parameters(REF * 30, REF, DTM, None)
    v.update_crf(3)
    if src_c:
        v.set_source_channels(src_c)
    return v


def run_frames(v, d_frames, stride, P_out, d_events, cap, d_off, per_frame=None):
    """NF single-frame launches (the form these rows run in) from a fresh state, after one untimed pass of the same
    (first-use allocations, module loads); returns ms per frame (CUDA events)."""
    for timed in (False, True):
        v.reset_state()
        v.update_crf(3)
        v.timer_start()
        for f in range(NF):
            v.integrate_frames_device(d_frames.ptr + f * stride, stride, 1, float(REF), d_events.ptr, cap, d_off.ptr)
            if per_frame:
                per_frame(f)
        ms = v.timer_stop()
        v.sync()
    return ms / NF


print(f"# {W}x{H}, {NF} frames of uniform noise, crf 3, one frame per launch; measured HBM peak {PEAK} GB/s")

# ---- row 1: raw serialisation (RGB: 11-byte wire records) ----------------------------------------------------
v = video(3)
P = W * H * 3
d_frames = v.device_alloc(P * NF)
v.synth_frames(d_frames, 0, NF, 1, 0xADDE5)
cap = P * 2
d_events = v.device_alloc(cap * 12)
d_off = v.device_alloc((v.n_chunks + 1) * 4)
d_raw = v.device_alloc(cap * 11)
base = run_frames(v, d_frames, P, P, d_events, cap, d_off)
with_raw = run_frames(v, d_frames, P, P, d_events, cap, d_off,
                      per_frame=lambda f: v.raw_encode_device(d_events.ptr, d_off.ptr + v.n_chunks * 4, cap, d_raw.ptr))
n_ev = int(d_off.to_host(np.uint32)[-1])
raw_us = (with_raw - base) * 1e3
print(f"row 1 raw_encode_kernel: {raw_us:7.1f} us/frame for {n_ev} events -> {(12 + 11) * n_ev / raw_us / 1e3:6.0f} GB/s "
      f"({(12 + 11) * n_ev / raw_us / 1e3 / PEAK:.2f} of peak; 12 B read + 11 B written per event); transcode alone {base * 1e3:.1f} us/frame")
if not quick:
    ev = d_events.to_host(A.EVENT_DTYPE, nbytes=n_ev * 12)
    t0 = time.perf_counter(); body = O.raw_encode(ev, 3); dt = time.perf_counter() - t0
    assert body == d_raw.to_host(nbytes=n_ev * 11).tobytes()
    print(f"      CPU port (1 thread, as the reference's serial ingest_event loop): {dt * 1e6:9.0f} us/frame -> {n_ev / dt / 1e6:.1f} Mevents/s")
for b in (d_frames, d_events, d_off, d_raw):
    b.free()
del v

# ---- row 2: RGB -> gray in front of a gray transcode -----------------------------------------------------------
# the same gray frames twice: converted on the device from RGB, and handed over already gray (converted by the oracle)
Pg = W * H
vc = video(1, src_c=3)
v3 = video(3)  # the generator writes W*H*C bytes per frame of the handle it is called on
d_rgb = v3.device_alloc(Pg * 3 * NF)
v3.synth_frames(d_rgb, 0, NF, 1, 0xADDE5)
v3.sync()
capg = Pg * 2
d_ev2 = vc.device_alloc(capg * 12)
d_of = vc.device_alloc((vc.n_chunks + 1) * 4)
from_rgb = run_frames(vc, d_rgb, Pg * 3, Pg, d_ev2, capg, d_of)
rgb_host = d_rgb.to_host().reshape(NF, H, W, 3)
t0 = time.perf_counter()
gray_host = np.stack([O.handle_color(rgb_host[f]) for f in range(NF)])
cpu_gray = (time.perf_counter() - t0) / NF
assert np.array_equal(gray_host[-1].reshape(-1), vc.input_frame().reshape(-1))
vg = video(1)
d_gray = vg.device_alloc(Pg * NF)
d_gray.from_host(gray_host)
d_ev = vg.device_alloc(capg * 12)
gray_only = run_frames(vg, d_gray, Pg, Pg, d_ev, capg, d_of)
g_us = (from_rgb - gray_only) * 1e3
print(f"row 2 rgb_to_gray_kernel: {g_us:6.1f} us/frame more than the same frames handed over gray (kernel + one more launch per frame; 5.0 us in the "
      f"ncu launch list: 8 MB per launch, launch-bound) -> {4 * Pg / g_us / 1e3:6.0f} GB/s; gray transcode alone {gray_only * 1e3:.1f} us/frame")
print(f"      CPU port (1 thread, f64 per px as utils/cv.rs:215-232): {cpu_gray * 1e6:9.0f} us/frame")
# the same as one call over all frames: one conversion launch + one integrate launch
for vv in (vc, vg):
    vv.reset_state(); vv.update_crf(3)
d_evb = vc.device_alloc(capg * 12 * NF)
t = []
for vv, buf, st in ((vc, d_rgb, Pg * 3), (vg, d_gray, Pg)):
    for rep in range(2):  # the first pass sizes the scratch run of gray frames
        vv.reset_state(); vv.update_crf(3)
        vv.integrate_frames_device(buf.ptr, st, 1, float(REF), d_evb.ptr, capg, None)  # the first frame of a fresh state goes alone
        vv.timer_start()
        vv.integrate_frames_device(buf.ptr + st, st, NF - 1, float(REF), d_evb.ptr, capg, None)
        ms = vv.timer_stop() / (NF - 1); vv.sync()
    t.append(ms)
print(f"      all frames in one call: {t[0] * 1e3:6.1f} us/frame from RGB, {t[1] * 1e3:6.1f} us/frame from gray -> {(t[0] - t[1]) * 1e3:5.1f} us/frame for the conversion "
      f"= {4 * Pg / ((t[0] - t[1]) * 1e3) / 1e3:6.0f} GB/s ({4 * Pg / ((t[0] - t[1]) * 1e3) / 1e3 / PEAK:.2f} of peak)")
d_evb.free()
for b in (d_rgb, d_ev2):
    b.free()
del vc

# ---- row 4: feature pass (FAST 9_16 per fired pixel + c_thresh radius reset) ------------------------------------
vg.update_detect_features(True, True)
with_feat = run_frames(vg, d_gray, Pg, Pg, d_ev, capg, d_of)
n_evg = int(d_of.to_host(np.uint32)[-1])
f_us = (with_feat - gray_only) * 1e3
print(f"row 4 feature pass (feature_kernel + c_thresh reset by dilation; noise: a feature at every seventh pixel): {f_us:6.1f} us/frame for {n_evg} events of a gray frame ({len(vg.new_features())} new features in the last frame)")
vg.update_detect_features(False, False)
if not quick:
    ovf = O.Video(W, H, 1, O.MODE_FRAME_PERFECT)
    ovf.time_parameters(REF * 30, REF, DTM, None); ovf.update_crf(3)
    nt = O.max_threads()
    tt = []
    for on in (False, True):
        ovf.update_detect_features(on, on)
        t0 = time.perf_counter()
        for f in range(2, 5):
            ovf.integrate_matrix(gray_host[f], float(REF), nt)
        tt.append((time.perf_counter() - t0) / 3)
    print(f"      CPU port, integrate_matrix with the feature pass minus without ({nt} threads for the pixels, the pass itself serial as in the reference): {(tt[1] - tt[0]) * 1e6:9.0f} us/frame")

# ---- row 3: INSTANTANEOUS framer in lock step with the transcoder, events never leave HBM -----------------------
vg.reset_state(); vg.update_crf(3)
fr = A.Framer(W, H, 1, 1, 3, A.TIME_ABSOLUTE_T, REF * 30, REF, DTM, output_fps=30.0, ring_frames=160)
ing = wr = 0.0
n_frames_out = 0
n_events_in = 0
for f in range(NF):
    vg.integrate_frames_device(d_gray.ptr + f * Pg, Pg, 1, float(REF), d_ev.ptr, capg, d_of.ptr)
    vg.sync()
    n_events_in += int(d_of.to_host(np.uint32)[-1])
    t0 = time.perf_counter(); ready = fr.ingest_events_device(d_ev.ptr, d_of.ptr); ing += time.perf_counter() - t0
    if ready:
        t0 = time.perf_counter(); out = fr.write_multi_frame_bytes(); wr += time.perf_counter() - t0
        n_frames_out += len(out)
t0 = time.perf_counter(); fr.flush_frame_buffer(); out = fr.write_multi_frame_bytes(); wr += time.perf_counter() - t0
n_frames_out += len(out)
print(f"row 3 framer: ingest_events_device {ing / NF * 1e6:7.1f} us/frame (host clock, sync included; {n_events_in / NF:.0f} events/frame -> "
      f"{n_events_in / ing / 1e6:.0f} Mevents/s), write_multi_frame_bytes {wr / max(n_frames_out, 1) * 1e6:7.1f} us per output frame ({n_frames_out} frames, D2H included)")
# the same frames' events ingested without a wait per call (adder_b200_framer_ingest_events_device_async): one buffer per frame
vg.reset_state(); vg.update_crf(3)
fr2 = A.Framer(W, H, 1, 1, 3, A.TIME_ABSOLUTE_T, REF * 30, REF, DTM, output_fps=30.0, ring_frames=160)
d_evs = [vg.device_alloc(capg * 12) for _ in range(NF)]
d_ofs = [vg.device_alloc((vg.n_chunks + 1) * 4) for _ in range(NF)]
for f in range(NF):
    vg.integrate_frames_device(d_gray.ptr + f * Pg, Pg, 1, float(REF), d_evs[f].ptr, capg, d_ofs[f].ptr)
vg.sync()
for rep in range(2):
    t0 = time.perf_counter()
    for f in range(NF if rep else 2):
        fr2.ingest_events_device_async(d_evs[f].ptr, d_ofs[f].ptr)
    ready = fr2.frame_ready()
    dt_async = time.perf_counter() - t0
    if not rep:  # a fresh framer for the timed pass
        fr2 = A.Framer(W, H, 1, 1, 3, A.TIME_ABSOLUTE_T, REF * 30, REF, DTM, output_fps=30.0, ring_frames=160)
ev_bytes = 12 * n_events_in / NF
st_bytes = 2 * 16 * n_events_in / NF  # one 16-byte state record read and written per run of events (about one run per event here)
print(f"      asynchronous ingest of {NF} frames, one wait at the end: {dt_async / NF * 1e6:7.1f} us/frame -> {(ev_bytes + st_bytes) / (dt_async / NF) / 1e9:6.0f} GB/s of records + pixel state "
      f"({(ev_bytes + st_bytes) / (dt_async / NF) / 1e9 / PEAK:.2f} of the HBM roofline; the ring bytes it touches are not counted)")
for b in d_evs + d_ofs:
    b.free()
if not quick:
    of = O.Framer(W, H, 1, 1, 3, O.TIME_ABSOLUTE_T, REF * 30, REF, DTM, output_fps=30.0)
    ov = O.Video(W, H, 1, O.MODE_FRAME_PERFECT)
    ov.time_parameters(REF * 30, REF, DTM, None); ov.update_crf(3)
    gray = gray_host.reshape(NF, H, W, 1)
    tc = 0.0
    nt = O.max_threads()
    for f in range(min(NF, 8)):
        eo, co = ov.integrate_matrix(gray[f], float(REF), nt)
        t0 = time.perf_counter(); of.ingest_events_events(eo, co); tc += time.perf_counter() - t0
    print(f"      CPU port, ingest_events_events (1 thread): {tc / min(NF, 8) * 1e6:9.0f} us/frame")
