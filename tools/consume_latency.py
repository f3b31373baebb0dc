#!/usr/bin/env python
"""Latency of the drop-in call itself: adder_b200_video_integrate_matrix (one frame, host buffers; what replaces the body of
Video::integrate_matrix behind Framed::consume) at 1080p RGB noise, with the caller's event buffer (a) pageable, (b) page-locked
through adder_b200_host_alloc, next to the oracle port of the same call on this box's host threads."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adder_codec_rs_b200 as A  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import synth  # noqa: E402

W, H, C, NF, REF = 1920, 1080, 3, 16, 255
frames = synth.frames(synth.NOISE, 0xADDE5, 0, NF, W, H, C)
P = W * H * C
for name, ev, fr in (("pageable frame and event buffer", np.empty(P * 2, dtype=A.EVENT_DTYPE), frames),
                     ("page-locked frame and event buffer (adder_b200_host_alloc)", np.asarray(A.pinned_empty((P * 2,), A.EVENT_DTYPE)), None)):
    if fr is None:
        fr = np.asarray(A.pinned_empty(frames.shape, np.uint8))
        fr[...] = frames
    ev.view(np.uint8)[:] = 0  # touch every page once
    v = A.Video(W, H, C)
    v.time_parameters(REF * 30, REF, 7650, None)
    v.update_crf(3)
    ts = []
    for f in range(NF):
        t0 = time.perf_counter()
        e, c = v.integrate_matrix(fr[f], float(REF), ev)
        ts.append(time.perf_counter() - t0)
    med = np.median(ts[3:])
    print(f"integrate_matrix, {name}: {med * 1e3:6.2f} ms per call ({len(e)} events, {len(e) * 12 / 1e6:.0f} MB out, {P / 1e6:.1f} MB in) -> {P / med / 1e6:7.0f} Mpx/s")
ov = O.Video(W, H, C, O.MODE_FRAME_PERFECT)
ov.time_parameters(REF * 30, REF, 7650, None)
ov.update_crf(3)
nt = O.max_threads()
ts = []
for f in range(8):
    t0 = time.perf_counter()
    ov.integrate_matrix(frames[f], float(REF), nt)
    ts.append(time.perf_counter() - t0)
med = np.median(ts[3:])
print(f"oracle port of the same call, {nt} threads: {med * 1e3:6.2f} ms per call -> {P / med / 1e6:7.0f} Mpx/s")
