"""Parity at BASELINE.json's full plane sizes (GPU box only).

Direct: a few frames of every config family at its real geometry against the multi-threaded oracle,
event for event (the oracle does ~80 Mpx/s, so a handful of full frames costs seconds).
Properties that do not need the oracle, over longer runs: raster order inside every frame, chunk
offsets that bracket exactly the chunk's rows, the stream of two row bands concatenated == the stream
of the whole plane, and bit-identical reruns from a reset state.
"""
import numpy as np
import pytest

import adder_codec_rs_b200 as A
from adder_codec_rs_b200 import sharding as S
from oracle import oracle_py as O
from tests import cases, synth

pytestmark = pytest.mark.gpu

# (name, w, h, c, kind, frames, configure kwargs): SURVEY.md §8(d) table, exact parameters
FULL = [
    ("cfg1_640x480_gradient", 640, 480, 1, synth.GRADIENT, 30, dict(dtm=255)),
    ("cfg2_1080p_rgb_noise_crf3", 1920, 1080, 3, synth.NOISE, 6, dict(crf=3)),
    ("cfg3_4k_gray_jitter_c0", 3840, 2160, 1, synth.JITTER, 5, dict(manual=(0, 0, 30, 1))),
    ("cfg3_4k_gray_jitter_c10", 3840, 2160, 1, synth.JITTER, 5, dict(manual=(10, 10, 30, 1))),
    ("cfg4_4k_rgb_noise_crf3", 3840, 2160, 3, synth.NOISE, 3, dict(crf=3)),
    ("cfg5_8k_gray_static_dtm2e20", 7680, 4320, 1, synth.STATIC_BLIPS, 3, dict(crf=3, ref=256, dtm=1 << 20)),
]


def _case(name, w, h, c, kind, n, kw):
    return cases.Case(name, w, h, c, kind, n, **kw)


@pytest.mark.parametrize("spec", FULL, ids=lambda s: s[0])
def test_full_size_frames_match_the_oracle(spec):
    case = _case(*spec)
    gv = A.Video(case.w, case.h, case.c)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    P = case.w * case.h * case.c
    d_frames = gv.device_alloc(P)
    cap = P * 3
    d_events = gv.device_alloc(cap * 12)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4)
    nt = O.max_threads()
    for f in range(case.n_frames):
        gv.synth_frames(d_frames, f, 1, case.kind, case.seed)
        gv.integrate_frames_device(d_frames.ptr, P, 1, case.time, d_events.ptr, cap, d_off.ptr)
        gv.sync()
        frame = d_frames.to_host().reshape(case.h, case.w, case.c)  # the device generator is checked against numpy elsewhere
        eo, co = ov.integrate_matrix(frame, case.time, nt)
        off = d_off.to_host(np.uint32)
        assert int(off[-1]) == len(eo), f"frame {f}: {int(off[-1])} vs {len(eo)} events"
        assert np.array_equal(np.diff(off), co), f"frame {f}: chunk lengths differ"
        eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}: event streams differ"
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    for b in (d_frames, d_events, d_off):
        b.free()


def _run_device(v, case, n_frames, on_frame):
    P = v.w * v.h * v.c
    d_frames = v.device_alloc(P)
    cap = P * 3
    d_events = v.device_alloc(cap * 12)
    d_off = v.device_alloc((v.n_chunks + 1) * 4)
    for f in range(n_frames):
        v.synth_frames(d_frames, f, 1, case.kind, case.seed)
        v.integrate_frames_device(d_frames.ptr, P, 1, case.time, d_events.ptr, cap, d_off.ptr)
        v.sync()
        off = d_off.to_host(np.uint32)
        on_frame(f, d_events.to_host(A.EVENT_DTYPE, nbytes=int(off[-1]) * 12), off)
    for b in (d_frames, d_events, d_off):
        b.free()


def test_raster_order_and_chunk_brackets_over_a_long_run():
    """1080p RGB noise, 40 frames: inside a frame (y, x, c) never decreases, a pixel's events are
    contiguous, and chunk k's records are exactly those with y == k (chunk_rows 1)."""
    case = _case("order_1080p", 1920, 1080, 3, synth.NOISE, 40, dict(crf=3))
    gv = A.Video(case.w, case.h, case.c)
    cases.configure(gv, case)
    sums = []

    def check(f, ev, off):
        key = (ev["y"].astype(np.int64) * case.w + ev["x"]) * case.c + ev["c"]
        assert np.all(np.diff(key) >= 0), f"frame {f}: not in raster order"
        assert np.array_equal(np.searchsorted(ev["y"], np.arange(case.h + 1)), off.astype(np.int64)), f"frame {f}: chunk offsets"
        assert np.all(ev["reserved"] == 0) and np.all(ev["c"] < 3)
        sums.append((len(ev), int(ev["t"].astype(np.uint64).sum()), int(ev["d"].astype(np.uint64).sum())))

    _run_device(gv, case, case.n_frames, check)
    # bit-identical rerun from a reset state (checksum of checksums)
    first = list(sums)
    sums.clear()
    gv.reset_state()
    cases.configure(gv, case)
    _run_device(gv, case, case.n_frames, check)
    assert sums == first
    assert gv.events_emitted() == 2 * sum(s[0] for s in first)


def test_two_bands_equal_the_whole_plane_at_4k():
    """4K gray jitter c=5, 24 frames: rows split into two bands (both on this GPU), streams concatenated
    in band order == the stream of the undivided plane.  No oracle involved."""
    case = _case("bands_4k", 3840, 2160, 1, synth.JITTER, 24, dict(manual=(5, 5, 30, 1)))
    whole = A.Video(case.w, case.h, case.c)
    cases.configure(whole, case)
    bands = [S.BandedVideo(case.w, case.h, case.c, r, 2, device=0) for r in range(2)]
    for b in bands:
        cases.configure(b, case)
    P = case.w * case.h
    d_frames = whole.device_alloc(P)
    cap = P * 2
    d_ev = whole.device_alloc(cap * 12)
    d_off = whole.device_alloc((whole.n_chunks + 1) * 4)
    d_evb = [b.device_alloc(cap * 12 // 2) for b in bands]
    d_offb = [b.device_alloc((b.n_chunks + 1) * 4) for b in bands]
    for f in range(case.n_frames):
        whole.synth_frames(d_frames, f, 1, case.kind, case.seed)
        whole.sync()
        whole.integrate_frames_device(d_frames.ptr, P, 1, case.time, d_ev.ptr, cap, d_off.ptr)
        whole.sync()
        n = int(d_off.to_host(np.uint32)[-1])
        want = d_ev.to_host(A.EVENT_DTYPE, nbytes=n * 12)
        got = []
        for b, de, do in zip(bands, d_evb, d_offb):
            pb = case.w * b.rows
            b.integrate_frames_device(d_frames.ptr + b.row0 * case.w, pb, 1, case.time, de.ptr, cap // 2, do.ptr)
            b.sync()
            nb = int(do.to_host(np.uint32)[-1])
            got.append(de.to_host(A.EVENT_DTYPE, nbytes=nb * 12))
        assert np.concatenate(got).tobytes() == want.tobytes(), f"frame {f}"


MULTI = [
    ("cfg2_1080p_rgb_noise_24_frames", 1920, 1080, 3, synth.NOISE, 24, dict(crf=3)),
    ("cfg3_4k_gray_jitter_c10_16_frames", 3840, 2160, 1, synth.JITTER, 16, dict(manual=(10, 10, 30, 1))),
    ("cfg5_8k_gray_static_8_frames", 7680, 4320, 1, synth.STATIC_BLIPS, 8, dict(crf=3, ref=256, dtm=1 << 20)),
]


@pytest.mark.parametrize("spec", MULTI, ids=lambda s: s[0])
def test_full_size_multi_frame_launch_matches_the_oracle(spec):
    """All frames through ONE launch (what bench.py times): thousands of tiles per frame, frames overlapping in
    flight, every frame's stream compared with the oracle run frame by frame."""
    case = _case(*spec)
    gv = A.Video(case.w, case.h, case.c)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    P = case.w * case.h * case.c
    nf = case.n_frames
    d_frames = gv.device_alloc(P * nf)
    gv.synth_frames(d_frames, 0, nf, case.kind, case.seed)
    cap = P * 2
    d_events = gv.device_alloc(cap * 12 * nf)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4 * nf)
    gv.integrate_frames_device(d_frames.ptr, P, nf, case.time, d_events.ptr, cap, d_off.ptr)
    gv.sync()
    offs = d_off.to_host(np.uint32).reshape(nf, gv.n_chunks + 1)
    nt = O.max_threads()
    for f in range(nf):
        frame = d_frames.to_host(nbytes=P, offset=f * P).reshape(case.h, case.w, case.c)
        eo, co = ov.integrate_matrix(frame, case.time, nt)
        assert int(offs[f, -1]) == len(eo), f"frame {f}: {int(offs[f, -1])} vs {len(eo)} events"
        assert np.array_equal(np.diff(offs[f]), co), f"frame {f}: chunk lengths differ"
        eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=f * cap * 12)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}: event streams differ"
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    for b in (d_frames, d_events, d_off):
        b.free()


def test_long_multi_frame_launch_at_bench_geometry():
    """The bench step in small: 1080p RGB noise, 120 frames through ONE launch.  Every frame's event total and chunk
    offsets against the oracle, the full streams of a few frames byte for byte, and the final display plane."""
    case = _case("bench_120", 1920, 1080, 3, synth.NOISE, 120, dict(crf=3))
    gv = A.Video(case.w, case.h, case.c)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    P = case.w * case.h * case.c
    nf = case.n_frames
    d_frames = gv.device_alloc(P * nf)
    gv.synth_frames(d_frames, 0, nf, case.kind, case.seed)
    cap = P * 2
    d_events = gv.device_alloc(cap * 12 * nf)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4 * nf)
    gv.integrate_frames_device(d_frames.ptr, P, 1, case.time, d_events.ptr, cap, d_off.ptr)  # a fresh state: the first frame goes alone
    gv.integrate_frames_device(d_frames.ptr + P, P, nf - 1, case.time, d_events.ptr + cap * 12, cap, d_off.ptr + (gv.n_chunks + 1) * 4)
    gv.sync()
    offs = d_off.to_host(np.uint32).reshape(nf, gv.n_chunks + 1)
    nt = O.max_threads()
    for f in range(nf):
        frame = d_frames.to_host(nbytes=P, offset=f * P).reshape(case.h, case.w, case.c)
        eo, co = ov.integrate_matrix(frame, case.time, nt)
        assert int(offs[f, -1]) == len(eo), f"frame {f}: {int(offs[f, -1])} vs {len(eo)} events"
        assert np.array_equal(np.diff(offs[f]), co), f"frame {f}: chunk lengths differ"
        if f in (1, 2, 57, 118, 119):
            eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=f * cap * 12)
            assert eg.tobytes() == eo.tobytes(), f"frame {f}: event streams differ"
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    for b in (d_frames, d_events, d_off):
        b.free()
