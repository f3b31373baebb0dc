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


# Frames whose complete streams are compared byte for byte in the long runs below: the first frames, the frames
# around the first Δt_max pop (dtm = 30 ref: a pixel that never changed pops its root at frame 30,
# event_pixel_tree.rs:394-396) and the end of the crf-3 c_thresh ramp (2 -> 7, +1 every 7 frames: frame ~35,
# :402-412), and the last frames.  Every other frame is compared by event total and per-chunk lengths.
FULL_STREAM_FRAMES = (0, 1, 2, 29, 30, 31, 32, 33, 34, 35, 36, 37, 62, 63)


def _one_launch_against_the_oracle(case, nf, cap_per_px, full_frames):
    """All nf frames through ONE integrate launch (what bench.py times) against the oracle run frame by frame:
    event totals and chunk lengths for every frame, whole streams for `full_frames`, and the display plane at the end."""
    gv = A.Video(case.w, case.h, case.c)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    P = case.w * case.h * case.c
    d_frames = gv.device_alloc(P * nf)
    gv.synth_frames(d_frames, 0, nf, case.kind, case.seed)
    cap = int(P * cap_per_px)
    d_events = gv.device_alloc(cap * 12 * nf)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4 * nf)
    l0 = gv.launch_count
    gv.integrate_frames_device(d_frames.ptr, P, nf, case.time, d_events.ptr, cap, d_off.ptr)
    gv.sync()
    assert gv.launch_count - l0 == 1, "the frames were meant to go through one launch"
    offs = d_off.to_host(np.uint32).reshape(nf, gv.n_chunks + 1)
    nt = O.max_threads()
    compared = 0
    for f in range(nf):
        frame = d_frames.to_host(nbytes=P, offset=f * P).reshape(case.h, case.w, case.c)
        eo, co = ov.integrate_matrix(frame, case.time, nt)
        assert int(offs[f, -1]) == len(eo), f"frame {f}: {int(offs[f, -1])} vs {len(eo)} events"
        assert np.array_equal(np.diff(offs[f]), co), f"frame {f}: chunk lengths differ"
        if f in full_frames:
            eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=f * cap * 12)
            assert eg.tobytes() == eo.tobytes(), f"frame {f}: event streams differ"
            compared += 1
    assert compared >= min(8, nf)
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    for b in (d_frames, d_events, d_off):
        b.free()
    return ov


# every BASELINE geometry, 64 frames through one launch (cap_per_px bounds the events of one frame)
LONG = [
    ("cfg2_1080p_rgb_noise_64_frames", 1920, 1080, 3, synth.NOISE, 64, dict(crf=3), 2),
    ("cfg3_4k_gray_jitter_c10_64_frames", 3840, 2160, 1, synth.JITTER, 64, dict(manual=(10, 10, 30, 1)), 2),
    ("cfg4_4k_rgb_noise_64_frames", 3840, 2160, 3, synth.NOISE, 64, dict(crf=3), 1.25),
    ("cfg5_8k_gray_static_collapse_64_frames", 7680, 4320, 1, synth.STATIC_BLIPS, 64, dict(crf=3, ref=256, dtm=1 << 20), 0.5),
    ("cfg5_8k_gray_static_normal_64_frames", 7680, 4320, 1, synth.STATIC_BLIPS, 64,
     dict(crf=3, ref=256, dtm=1 << 20, multi_mode=O.MULTI_NORMAL), 0.5),
]


@pytest.mark.parametrize("spec", LONG, ids=lambda s: s[0])
def test_full_size_64_frames_in_one_launch_match_the_oracle(spec):
    """VERDICT r1 task 1(a): at every BASELINE geometry the run goes past the first Δt_max pop and the end of the
    c ramp, so pop_top_event, Collapse / D_EMPTY and the park arena are compared at full size."""
    case = _case(*spec[:7])
    _one_launch_against_the_oracle(case, case.n_frames, spec[7], FULL_STREAM_FRAMES)


@pytest.mark.parametrize("c", list(range(11)))
def test_cfg3_contrast_threshold_sweep_at_4k(c):
    """VERDICT r1 task 1(b): BASELINE configs[2], every c_thresh of the sweep 0..=10 (quality_manual(c, c, 30, 1, 0.0)),
    40 frames of 4K gray jitter through one launch."""
    case = _case(f"cfg3_4k_c{c}", 3840, 2160, 1, synth.JITTER, 40, dict(manual=(c, c, 30, 1)))
    _one_launch_against_the_oracle(case, 40, 2, (0, 1, 2, 29, 30, 31, 32, 33, 34, 35, 36, 39))


@pytest.mark.parametrize("multi", [O.MULTI_COLLAPSE, O.MULTI_NORMAL], ids=["collapse", "normal"])
def test_cfg5_long_integration_past_delta_t_max(multi):
    """VERDICT r1 task 1(c): ref 256, dtm 2^20 = 4096 ref.  A small static plane (every intensity 0..255 present, no
    blips) run for 4200 frames: the deep stacks (11+ live nodes) and their Δt_max pop at frame 4096 are compared with
    the oracle, every event of every frame, and the complete pixel state at the end."""
    w, h, nf = 64, 24, 4200
    base = synth.frame(synth.NOISE, 0xC0FFEE, 0, w, h, 1)
    base.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    gv = A.Video(w, h, 1)
    ov = O.Video(w, h, 1, O.MODE_FRAME_PERFECT)
    for v in (gv, ov):
        assert v.time_parameters(256 * 30, 256, 1 << 20, None)
        v.write_out(None, multi)
        v.update_crf(3)
    P = w * h
    batch = 300
    frames = np.broadcast_to(base[None], (batch, h, w, 1)).copy()
    d_frames = gv.device_alloc(P * batch)
    d_frames.from_host(frames)
    cap = P * 16
    d_events = gv.device_alloc(cap * 12 * batch)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4 * batch)
    max_len = 0
    total = 0
    for f0 in range(0, nf, batch):
        gv.integrate_frames_device(d_frames.ptr, P, batch, 256.0, d_events.ptr, cap, d_off.ptr)
        gv.sync()
        offs = d_off.to_host(np.uint32).reshape(batch, gv.n_chunks + 1)
        for k in range(batch):
            eo, co = ov.integrate_matrix(base, 256.0)
            assert int(offs[k, -1]) == len(eo), f"frame {f0 + k}: {int(offs[k, -1])} vs {len(eo)} events"
            assert np.array_equal(np.diff(offs[k]), co), f"frame {f0 + k}"
            if len(eo):
                eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=k * cap * 12)
                assert eg.tobytes() == eo.tobytes(), f"frame {f0 + k}: event streams differ"
            total += len(eo)
        if f0 + batch in (3900, 4200):
            for i in range(0, P, 7):
                a = cases.canonical_oracle_px(ov.px(i))
                b = cases.canonical(gv.px_dict(i))
                assert a == b, f"after frame {f0 + batch}, pixel {i}: state differs\noracle {a}\ngpu    {b}"
                max_len = max(max_len, a["length"])
    assert max_len >= 11, f"deepest stack seen: {max_len}"
    assert total >= P - 16, "the Δt_max pops at frame 4096 should have produced an event for (nearly) every pixel"
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())


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
