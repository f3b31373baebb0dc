"""One launch spanning several frames (adder_b200_video_integrate_frames_device with n_frames > 1), GPU box only.

Inside such a launch the tiles of frame f+1 start while those of frame f are still in flight, ordered through the
per-tile status words; events of a frame are parked in a per-CTA arena.  These tests force that machinery through
every tile shape, small planes (few tiles: the status ring wraps often), uneven call boundaries, bursts of events per
pixel and more frames than one launch may span, always against the oracle run frame by frame."""
import numpy as np
import pytest

import adder_codec_rs_b200 as A
from adder_codec_rs_b200 import binding as B
from oracle import oracle_py as O
from tests import cases, synth

pytestmark = pytest.mark.gpu


def _pair(case):
    gv = A.Video(case.w, case.h, case.c, A.MODE_FRAME_PERFECT)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    return gv, ov


def _run_calls(gv, ov, case, frames, splits, cap_per_px=4, before_call=None):
    """frames[splits[k]:splits[k+1]] per integrate_frames_device call; every frame checked against the oracle."""
    P = case.w * case.h * case.c
    nck = gv.n_chunks
    longest = max(b - a for a, b in zip(splits[:-1], splits[1:]))
    stride = P * cap_per_px
    d_frames = gv.device_alloc(P * longest)
    d_events = gv.device_alloc(stride * 12 * longest)
    d_off = gv.device_alloc((nck + 1) * 4 * longest)
    for a, b in zip(splits[:-1], splits[1:]):
        if before_call:
            before_call(a)
        nf = b - a
        d_frames.from_host(frames[a:b])
        gv.integrate_frames_device(d_frames.ptr, P, nf, case.time, d_events.ptr, stride, d_off.ptr)
        gv.sync()
        offs = d_off.to_host(np.uint32, nbytes=(nck + 1) * 4 * nf).reshape(nf, nck + 1)
        for f in range(nf):
            eo, co = ov.integrate_matrix(frames[a + f], case.time)
            assert offs[f, -1] == len(eo), f"frame {a + f}: {offs[f, -1]} vs {len(eo)} events"
            assert np.array_equal(np.diff(offs[f]), co), f"frame {a + f}: chunk lengths"
            eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=f * stride * 12)
            assert eg.tobytes() == eo.tobytes(), f"frame {a + f}: events"
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    for buf in (d_frames, d_events, d_off):
        buf.free()


def _state_equal(gv, ov, n, step):
    for i in range(0, n, step):
        a = cases.canonical_oracle_px(ov.px(i))
        b = cases.canonical(gv.px_dict(i))
        assert a == b, f"pixel {i}: state differs\noracle {a}\ngpu    {b}"


@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "cfg5_static_normal", "ragged_37x13x3_chunk4", "ragged_3x700_chunk64",
                                  "jitter_dtm4_collapse", "cfg1_gradient_dtm_eq_ref"])
def test_multi_frame_launch_every_tile_shape(name, r, monkeypatch):
    monkeypatch.setenv("ADDER_B200_R", str(r))
    case = cases.CASES_BY_NAME[name]
    gv, ov = _pair(case)
    frames = case.frames()
    n = case.n_frames
    _run_calls(gv, ov, case, frames, [0, 1, 1 + (n - 1) // 3, n])  # a single frame, a short launch, a long one
    P = case.w * case.h * case.c
    _state_equal(gv, ov, P, max(1, P // 150))


def test_more_frames_than_one_launch_spans():
    """700 frames in one call: the library cuts it into launches of at most 512 frames; the status ring of a plane
    with a single tile wraps hundreds of times."""
    case = cases.Case("long_run", 24, 9, 3, synth.JITTER, 700, manual=(3, 9, 30, 2))
    gv, ov = _pair(case)
    _run_calls(gv, ov, case, case.frames(), [0, 700])
    _state_equal(gv, ov, 24 * 9 * 3, 5)


def test_bursts_of_events_inside_a_multi_frame_launch():
    """Normal mode, deep stacks, then c_thresh drops to 0 over the plane between two calls: the first frames of the
    second launch pop up to 7 events per pixel (beyond the shared-memory slot: the per-CTA arena), while the next
    frames of the same launch are already in flight."""
    case = cases.Case("deep_pop_mf", 64, 32, 1, synth.JITTER, 170, manual=(25, 25, 4096, 1), ref=256, dtm=1 << 20,
                      multi_mode=O.MULTI_NORMAL)
    gv, ov = _pair(case)

    def before(a):
        if a == 150:
            gv.set_c_thresh_rect(0, 0, 63, 31, 0)
            ov.set_c_thresh_rect(0, 0, 63, 31, 0)

    _run_calls(gv, ov, case, case.frames(), [0, 150, 170], cap_per_px=8, before_call=before)
    _state_equal(gv, ov, 64 * 32, 7)


def test_capacity_overflow_in_a_later_frame_of_a_launch_is_reported():
    case = cases.CASES_BY_NAME["cfg2_rgb_noise_crf3"]
    gv, ov = _pair(case)
    frames = case.frames()
    P = case.w * case.h * case.c
    need = [len(ov.integrate_matrix(frames[f], case.time)[0]) for f in range(6)]
    assert need[0] == 0 and min(need[1:]) > 1  # nothing to give in the first frame, then about one event per pixel
    cap = max(need[1:]) - 1  # enough for some frames of the launch, one record short for at least one later frame
    d_frames = gv.device_alloc(P * 6)
    d_frames.from_host(frames[:6])
    d_events = gv.device_alloc(cap * 12 * 6)
    gv.integrate_frames_device(d_frames.ptr, P, 6, case.time, d_events.ptr, cap, None)
    with pytest.raises(A.AdderError) as e:
        gv.sync()
    assert e.value.code == B.ERR_CAPACITY


def test_band_offset_inside_a_multi_frame_launch():
    """A row band (set_row_offset) run as one launch over all frames emits the same records as the rows of the whole
    plane run frame by frame."""
    from adder_codec_rs_b200 import sharding as S

    case = cases.CASES_BY_NAME["cfg2_rgb_noise_crf3"]
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    frames = case.frames()
    band = S.BandedVideo(case.w, case.h, case.c, 1, 2, device=0)
    cases.configure(band, case)
    nf = case.n_frames
    pb = case.w * case.c * band.rows
    d_frames = band.device_alloc(pb * nf)
    d_frames.from_host(np.ascontiguousarray(frames[:, band.row0:band.row0 + band.rows]))
    stride = pb * 3
    d_events = band.device_alloc(stride * 12 * nf)
    d_off = band.device_alloc((band.n_chunks + 1) * 4 * nf)
    band.integrate_frames_device(d_frames.ptr, pb, nf, case.time, d_events.ptr, stride, d_off.ptr)
    band.sync()
    offs = d_off.to_host(np.uint32).reshape(nf, band.n_chunks + 1)
    for f in range(nf):
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        want = eo[(eo["y"] >= band.row0) & (eo["y"] < band.row0 + band.rows)]
        got = d_events.to_host(A.EVENT_DTYPE, nbytes=int(offs[f, -1]) * 12, offset=f * stride * 12)
        assert got.tobytes() == want.tobytes(), f"frame {f}"
