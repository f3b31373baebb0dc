"""Parity of the CUDA path against the oracle, through the C ABI (GPU box only)."""
import numpy as np
import pytest

import adder_codec_rs_b200 as A
from adder_codec_rs_b200 import binding as B
from oracle import oracle_py as O
from tests import cases, synth

pytestmark = pytest.mark.gpu


def _pair(case, **kw):
    gv = A.Video(case.w, case.h, case.c, A.MODE_FRAME_PERFECT, **kw)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    return gv, ov


def _assert_state_equal(gv, ov, n, step=1):
    for i in range(0, n, step):
        a = cases.canonical_oracle_px(ov.px(i))
        b = cases.canonical(gv.px_dict(i))
        assert a == b, f"pixel {i}: state differs\noracle {a}\ngpu    {b}"


@pytest.mark.parametrize("case", cases.CASES, ids=lambda c: c.name)
def test_events_chunks_display_state_match_oracle(case):
    """Frame by frame through integrate_matrix (host buffers): events bit-exact and in order,
    per-chunk lengths, running_intensities, in_interval_count, and the final per-pixel state."""
    gv, ov = _pair(case)
    frames = case.frames()
    total = 0
    for f in range(case.n_frames):
        if case.roi and case.roi[0] == f:
            gv.set_c_thresh_rect(*case.roi[1:])
            ov.set_c_thresh_rect(*case.roi[1:])
        eg, cg = gv.integrate_matrix(frames[f], case.time)
        eo, co = ov.integrate_matrix(frames[f], case.time)
        assert len(eg) == len(eo), f"frame {f}: {len(eg)} vs {len(eo)} events"
        assert eg.tobytes() == eo.tobytes(), f"frame {f}: event streams differ"
        assert np.array_equal(cg, co), f"frame {f}: chunk counts differ"
        assert np.array_equal(gv.running_intensities(), ov.running_intensities()), f"frame {f}: display bytes differ"
        total += len(eg)
    assert gv.in_interval_count == ov.in_interval_count
    assert gv.events_emitted() == total
    n = case.w * case.h * case.c
    _assert_state_equal(gv, ov, n, step=max(1, n // 400))


@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "jitter_dtm4_collapse", "ragged_37x13x3_chunk4", "cfg5_static_normal"])
def test_batched_host_form_equals_frame_by_frame(name):
    """integrate_frames_host (pipelined copies) == n calls of integrate_matrix."""
    case = cases.CASES_BY_NAME[name]
    gv, ov = _pair(case)
    frames = case.frames()
    exp, exp_fc, exp_cc = [], [], []
    for f in range(case.n_frames):
        eo, co = ov.integrate_matrix(frames[f], case.time)
        exp.append(eo)
        exp_fc.append(len(eo))
        exp_cc.append(co)
    exp = np.concatenate(exp)
    pinned_frames = A.pinned_empty(frames.shape, np.uint8)
    pinned_frames[...] = frames
    out = A.pinned_empty((len(exp) + 16,), A.EVENT_DTYPE)
    ev, fc, cc = gv.integrate_frames_host(np.asarray(pinned_frames), case.time, np.asarray(out))
    assert ev.tobytes() == exp.tobytes()
    assert np.array_equal(fc, np.array(exp_fc, dtype=np.uint64))
    assert np.array_equal(cc, np.stack(exp_cc))
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())


@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "cfg3_jitter_c5", "blips_dtm16_collapse"])
def test_device_resident_form(name):
    """Frames already in HBM (generated there by the synth kernel), events left in HBM."""
    case = cases.CASES_BY_NAME[name]
    gv, ov = _pair(case)
    P = case.w * case.h * case.c
    nf = case.n_frames
    d_frames = gv.device_alloc(P * nf)
    gv.synth_frames(d_frames, 0, nf, case.kind, case.seed)
    host_frames = case.frames()
    assert np.array_equal(d_frames.to_host().reshape(host_frames.shape), host_frames), "device and numpy generators differ"
    stride = P * 3  # records per frame
    d_events = gv.device_alloc(stride * nf * 12)
    nck = gv.n_chunks
    d_off = gv.device_alloc((nck + 1) * 4 * nf)
    gv.integrate_frames_device(d_frames.ptr, P, nf, case.time, d_events.ptr, stride, d_off.ptr)
    gv.sync()
    offs = d_off.to_host(np.uint32).reshape(nf, nck + 1)
    for f in range(nf):
        eo, co = ov.integrate_matrix(host_frames[f], case.time)
        assert offs[f, -1] == len(eo), f"frame {f}"
        assert np.array_equal(np.diff(offs[f]), co)
        eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=f * stride * 12)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}"
    for b in (d_frames, d_events, d_off):
        b.free()


def test_capacity_error_keeps_the_events():
    case = cases.CASES_BY_NAME["cfg2_rgb_noise_crf3"]
    gv, ov = _pair(case)
    frames = case.frames()
    for f in range(3):
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        small = np.empty(4, dtype=A.EVENT_DTYPE)
        n = B.C.c_uint64()
        rc = gv.L.adder_b200_video_integrate_matrix(gv.v, frames[f].ctypes.data, 0, case.time, small.ctypes.data, len(small), None, B.C.byref(n))
        if len(eo) > 4:
            assert rc == B.ERR_CAPACITY and n.value == len(eo)
            big = np.empty(n.value, dtype=A.EVENT_DTYPE)
            rc = gv.L.adder_b200_video_fetch_events(gv.v, big.ctypes.data, len(big), None, B.C.byref(n))
            assert rc == 0 and big.tobytes() == eo.tobytes()
        else:
            assert rc == 0


def test_device_capacity_overflow_is_reported():
    case = cases.CASES_BY_NAME["cfg2_rgb_noise_crf3"]
    gv, _ = _pair(case)
    P = case.w * case.h * case.c
    d_frames = gv.device_alloc(P * 4)
    gv.synth_frames(d_frames, 0, 4, case.kind, case.seed)
    d_events = gv.device_alloc(16 * 4 * 12)
    gv.integrate_frames_device(d_frames.ptr, P, 4, case.time, d_events.ptr, 16, None)
    with pytest.raises(A.AdderError) as e:
        gv.sync()
    assert e.value.code == B.ERR_CAPACITY


def test_arena_depth_overflow_is_reported():
    case = cases.CASES_BY_NAME["cfg5_static_collapse"]
    gv, _ = _pair(case, max_depth=3)
    frames = case.frames()
    with pytest.raises(A.AdderError) as e:
        for f in range(case.n_frames):
            gv.integrate_matrix(frames[f], case.time)
    assert e.value.code == B.ERR_ARENA_DEPTH


def test_continuous_mode_is_refused():
    gv = A.Video(8, 8, 1, A.MODE_CONTINUOUS)
    with pytest.raises(A.AdderError) as e:
        gv.integrate_matrix(np.zeros((8, 8, 1), np.uint8), 255.0)
    assert e.value.code == B.ERR_UNSUPPORTED


def test_bad_params():
    with pytest.raises(A.AdderError):
        A.Video(0, 8, 1)
    gv = A.Video(8, 8, 1)
    assert gv.time_parameters(100, 255, 100, None) is False  # dtm < ref: kept, like video.rs:518-523
    assert gv.info().delta_t_max == 7650
    with pytest.raises(A.AdderError):
        gv.chunk_rows(0)


def test_reset_state_equals_fresh_video():
    case = cases.CASES_BY_NAME["cfg3_jitter_c5"]
    gv, ov = _pair(case)
    frames = case.frames()
    for f in range(10):
        gv.integrate_matrix(frames[f], case.time)
    gv.reset_state()
    cases.configure(gv, case)
    for f in range(20):
        eg, _ = gv.integrate_matrix(frames[f], case.time)
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        assert eg.tobytes() == eo.tobytes()


def test_mid_size_noise_against_oracle():
    """A plane spanning thousands of tiles (look-back chains, many chunks): 640x360x3 noise, crf 3."""
    case = cases.Case("mid_noise", 640, 360, 3, synth.NOISE, 12, crf=3)
    gv, ov = _pair(case)
    frames = case.frames()
    nthreads = O.max_threads()
    for f in range(case.n_frames):
        eg, cg = gv.integrate_matrix(frames[f], case.time)
        eo, co = ov.integrate_matrix(frames[f], case.time, nthreads)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}"
        assert np.array_equal(cg, co)
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())


def test_framed_source_mirror():
    """Framed(...).crf().auto_time_parameters().write_out() ... consume(): the reference's usage
    (examples/framed_video_to_adder.rs:20-57) against the oracle driven the same way."""
    w, h = 48, 20
    rgb = synth.frames(synth.JITTER, 7, 0, 30, w, h, 3)
    src = A.Framed(list(rgb), w, h, color_input=False, source_fps=24.0)
    src = src.crf(0).frame_start(1).auto_time_parameters(255, 6120, None).write_out(A.TIME_DELTA_T, A.MULTI_NORMAL, (0, 0, 10))
    with pytest.raises(Exception):
        src.auto_time_parameters(255, 6121, None)
    ov = O.Video(w, h, 1, O.MODE_FRAME_PERFECT)
    ov.update_crf(0)
    assert ov.time_parameters(int(255 * 24), 255, 6120, None)
    ov.write_out(O.TIME_DELTA_T, O.MULTI_NORMAL)
    ov.set_crf_parameters(0, 0, 10)
    for f in range(1, 30):
        chunks = src.consume()
        assert len(chunks) == h  # chunk_rows 1 -> one Vec<Event> per row (driver.rs:566)
        eo, co = ov.integrate_matrix(O.handle_color(rgb[f]), 255.0)
        assert np.concatenate(chunks).tobytes() == eo.tobytes()
        assert [len(c) for c in chunks] == list(co)
    with pytest.raises(Exception):
        src.consume()


@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "cfg5_static_normal", "ragged_37x13x3_chunk4", "ragged_3x700_chunk64",
                                  "jitter_dtm4_normal", "cfg1_gradient_dtm_eq_ref", "ragged_300x2x2_chunk5"])
def test_every_tile_shape(name, r, monkeypatch):
    """The kernel is compiled for 1, 2, 4 and 8 sub-tiles per CTA (chosen by plane size): force each."""
    monkeypatch.setenv("ADDER_B200_R", str(r))
    case = cases.CASES_BY_NAME[name]
    gv, ov = _pair(case)
    frames = case.frames()
    for f in range(case.n_frames):
        eg, cg = gv.integrate_matrix(frames[f], case.time)
        eo, co = ov.integrate_matrix(frames[f], case.time)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}"
        assert np.array_equal(cg, co), f"frame {f}"
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    n = case.w * case.h * case.c
    _assert_state_equal(gv, ov, n, step=max(1, n // 200))


def test_many_events_per_pixel_go_through_the_park_arena():
    """Normal mode, long integration under a wide threshold, then the threshold drops to 0 over the
    whole plane (the ROI write): every pixel pops a deep stack in one frame — up to 7 events per pixel,
    more than the shared-memory slots hold, 3.7 events per pixel on average in that frame."""
    case = cases.Case("deep_pop", 64, 32, 1, synth.JITTER, 170, manual=(25, 25, 4096, 1), ref=256, dtm=1 << 20,
                      multi_mode=O.MULTI_NORMAL, roi=(150, 0, 0, 63, 31, 0))
    gv, ov = _pair(case)
    frames = case.frames()
    most = 0
    for f in range(case.n_frames):
        if case.roi[0] == f:
            gv.set_c_thresh_rect(*case.roi[1:])
            ov.set_c_thresh_rect(*case.roi[1:])
        eg, cg = gv.integrate_matrix(frames[f], case.time)
        eo, co = ov.integrate_matrix(frames[f], case.time)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}"
        assert np.array_equal(cg, co)
        if len(eo):
            most = max(most, np.bincount(eo["y"].astype(np.int64) * case.w + eo["x"]).max())
    assert most >= 6, most
    _assert_state_equal(gv, ov, case.w * case.h, step=7)


@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "cfg3_jitter_c5", "ragged_37x13x3_chunk4", "cfg5_static_normal"])
def test_raw_stream_body_equals_the_oracle_encoder(name):
    """integrate_frames_host_raw: the wire bytes RawOutput would write for the same events
    (9-byte records for one channel, 11-byte for colour), and the header for this video."""
    case = cases.CASES_BY_NAME[name]
    gv, ov = _pair(case)
    frames = case.frames()
    want = b""
    exp_fc = []
    for f in range(case.n_frames):
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        want += O.raw_encode(eo, case.c)
        exp_fc.append(len(eo))
    pinned = A.pinned_empty(frames.shape, np.uint8)
    pinned[...] = frames
    out = A.pinned_empty((len(want) + 64,), np.uint8)
    body, fc, _ = gv.integrate_frames_host_raw(np.asarray(pinned), case.time, np.asarray(out))
    assert gv.raw_event_size == (9 if case.c == 1 else 11)
    assert body.tobytes() == want
    assert np.array_equal(fc, np.array(exp_fc, dtype=np.uint64))
    i = gv.info()
    assert gv.raw_header() == O.raw_header(case.w, case.h, case.c, i.tps, i.ref_time, i.delta_t_max, version=3, time_mode=i.time_mode)
    assert gv.raw_header(version=1) == O.raw_header(case.w, case.h, case.c, i.tps, i.ref_time, i.delta_t_max, version=1)
    assert A.raw_eof() == O.raw_eof()


def test_raw_encode_device_resident_events():
    """Events left in HBM by the device-resident form, serialised in HBM, count read from the chunk offsets."""
    case = cases.Case("mid_noise_gray", 320, 200, 1, synth.NOISE, 6, crf=3)
    gv, ov = _pair(case)
    P = case.w * case.h * case.c
    frames = case.frames()
    d_frames = gv.device_alloc(P)
    stride = P * 3
    d_events = gv.device_alloc(stride * 12)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4)
    d_raw = gv.device_alloc(stride * 9)
    for f in range(case.n_frames):
        d_frames.from_host(frames[f])
        gv.integrate_frames_device(d_frames.ptr, P, 1, case.time, d_events.ptr, stride, d_off.ptr)
        gv.raw_encode_device(d_events.ptr, d_off.ptr + gv.n_chunks * 4, stride, d_raw.ptr)
        gv.sync()
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        assert d_raw.to_host(nbytes=len(eo) * 9).tobytes() == O.raw_encode(eo, 1), f"frame {f}"


def test_colour_source_gray_transcode_on_device():
    """set_source_channels(3): handle_color (utils/cv.rs:215-232) on the device in front of the integrate
    kernel, through all three forms, against the oracle's handle_color + transcode."""
    w, h, nf = 52, 21, 17  # odd sizes: the 4-pixel groups of the conversion kernel end ragged
    rgb = synth.frames(synth.NOISE, 11, 0, nf, w, h, 3)
    rgb[1] = 255
    rgb[2] = 0
    gv = A.Video(w, h, 1)
    ov = O.Video(w, h, 1, O.MODE_FRAME_PERFECT)
    for v in (gv, ov):
        v.update_crf(2)
    gv.set_source_channels(3)
    exp = []
    for f in range(nf):
        gray = O.handle_color(rgb[f])
        eo, co = ov.integrate_matrix(gray, 255.0)
        exp.append(eo)
        if f < 4:  # single-frame host form
            eg, cg = gv.integrate_matrix(rgb[f], 255.0)
            assert eg.tobytes() == eo.tobytes() and np.array_equal(cg, co), f"frame {f}"
            assert np.array_equal(gv.input_frame(), gray)
    # batched host form
    pinned = A.pinned_empty((4, h, w, 3), np.uint8)
    pinned[...] = rgb[4:8]
    out = A.pinned_empty((sum(len(e) for e in exp[4:8]) + 8,), A.EVENT_DTYPE)
    ev, _, _ = gv.integrate_frames_host(np.asarray(pinned), 255.0, np.asarray(out))
    assert ev.tobytes() == np.concatenate(exp[4:8]).tobytes()
    # device-resident form
    d_rgb = gv.device_alloc(4 * h * w * 3)
    d_rgb.from_host(rgb[8:12])
    stride = w * h * 3
    d_events = gv.device_alloc(4 * stride * 12)
    d_off = gv.device_alloc(4 * (gv.n_chunks + 1) * 4)
    gv.integrate_frames_device(d_rgb.ptr, 0, 4, 255.0, d_events.ptr, stride, d_off.ptr)
    gv.sync()
    offs = d_off.to_host(np.uint32).reshape(4, -1)
    for k in range(4):
        n = int(offs[k, -1])
        assert n == len(exp[8 + k])
        assert d_events.to_host(A.EVENT_DTYPE, nbytes=n * 12, offset=k * stride * 12).tobytes() == exp[8 + k].tobytes()
    assert np.array_equal(gv.input_frame(), O.handle_color(rgb[11]))
    # the same with padded frames (frame_stride > H*W*3): frame-by-frame conversion into the scratch run, one integrate launch
    pad = stride + 20
    d_pad = gv.device_alloc(5 * pad)
    for k in range(5):
        d_pad.from_host(rgb[12 + k], offset=k * pad)
    d_events5 = gv.device_alloc(5 * stride * 12)
    d_off5 = gv.device_alloc(5 * (gv.n_chunks + 1) * 4)
    gv.integrate_frames_device(d_pad.ptr, pad, 5, 255.0, d_events5.ptr, stride, d_off5.ptr)
    gv.sync()
    offs = d_off5.to_host(np.uint32).reshape(5, -1)
    for k in range(5):
        n = int(offs[k, -1])
        assert n == len(exp[12 + k])
        assert d_events5.to_host(A.EVENT_DTYPE, nbytes=n * 12, offset=k * stride * 12).tobytes() == exp[12 + k].tobytes()
    assert np.array_equal(gv.input_frame(), O.handle_color(rgb[16]))
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())


def test_raw_adder_file_layout(tmp_path):
    """Header + body + EOF written through RawAdderWriter parse back to the oracle's events
    (adder-info derives the event count as (eof_pos - 1 - header_size) / event_size, adder-info/src/main.rs:42)."""
    import struct

    from adder_codec_rs_b200.framed import RawAdderWriter

    case = cases.CASES_BY_NAME["cfg3_jitter_c5"]
    gv, ov = _pair(case)
    frames = case.frames()
    exp = np.concatenate([ov.integrate_matrix(frames[f], case.time)[0] for f in range(case.n_frames)])
    pinned = A.pinned_empty(frames.shape, np.uint8)
    pinned[...] = frames
    out = A.pinned_empty((len(exp) * 9 + 64,), np.uint8)
    path = tmp_path / "out.adder"
    with open(path, "wb") as fh:
        wr = RawAdderWriter(fh, gv)
        body, _, _ = gv.integrate_frames_host_raw(np.asarray(pinned), case.time, np.asarray(out))
        wr.write_body(body)
        wr.close()
    raw = open(path, "rb").read()
    assert raw[:5] == b"adder" and raw[5] == 3 and len(raw) == 37 + 9 * len(exp) + 11
    w, h, tps, ref, dtm = struct.unpack(">HHIII", raw[7:23])
    assert (w, h, ref, dtm, raw[23], raw[24]) == (case.w, case.h, case.ref, case.dtm, 9, 1)
    rec = np.frombuffer(raw[37:-11], dtype=np.dtype([("x", ">u2"), ("y", ">u2"), ("d", "u1"), ("t", ">u4")]))
    assert np.array_equal(rec["x"], exp["x"]) and np.array_equal(rec["y"], exp["y"])
    assert np.array_equal(rec["d"], exp["d"]) and np.array_equal(rec["t"], exp["t"])
    assert raw[-11:] == O.raw_eof() and wr.n_events == len(exp)


@pytest.mark.parametrize("c,adjust,chunk_rows,reset", [(1, True, 1, None), (3, True, 1, None), (1, False, 5, None), (1, True, 64, None),
                                                       (1, True, 1, "dilate"), (3, True, 4, "dilate"), (1, True, 1, "list"),
                                                       (3, True, 1, "noise")])
def test_feature_detection_pass_matches_the_oracle(c, adjust, chunk_rows, reset, monkeypatch):
    """handle_features (video.rs:883-1113) at the end of integrate_matrix: feature sets, newly found features
    and — with rate adjustment — the c_thresh reset around them, which changes the following frames' events.
    The reset has two device forms (per feature / dilation of the new-feature bitmap, chosen from the count):
    both are forced here, and the noise case (a feature at every seventh pixel) takes the dilation by itself."""
    w, h, nf = 96, 64, 40
    if reset in ("dilate", "list"):
        monkeypatch.setenv("ADDER_B200_FEATURE_RESET", reset)
    frames = synth.moving_blocks(3, nf, w, h, c)
    if reset == "noise":
        nf = 12
        frames = cases.Case("feat_noise", w, h, c, synth.NOISE, nf, crf=6).frames()
    gv = A.Video(w, h, c)
    ov = O.Video(w, h, c, O.MODE_FRAME_PERFECT)
    n_new = 0
    for v in (gv, ov):
        if chunk_rows != 1:
            v.chunk_rows(chunk_rows)
        v.update_crf(6)  # baseline 7, max 13, feature radius = 64/25 = 2
        v.update_detect_features(True, adjust)
    for f in range(nf):
        eg, cg = gv.integrate_matrix(frames[f], 255.0)
        eo, co = ov.integrate_matrix(frames[f], 255.0)
        assert eg.tobytes() == eo.tobytes(), f"frame {f}"
        assert np.array_equal(cg, co)
        assert np.array_equal(gv.new_features(), ov.new_features()), f"frame {f}: new features differ"
        assert np.array_equal(gv.feature_mask(), ov.feature_mask()), f"frame {f}: feature sets differ"
        n_new += len(ov.new_features())
    assert n_new > 20, n_new
    _assert_state_equal(gv, ov, w * h * c, step=3)
    # switching detection off stops the pass; the sets stay as they are
    gv.update_detect_features(False)
    ov.update_detect_features(False)
    eg, _ = gv.integrate_matrix(frames[0], 255.0)
    eo, _ = ov.integrate_matrix(frames[0], 255.0)
    assert eg.tobytes() == eo.tobytes() and len(gv.new_features()) == 0 == len(ov.new_features())


def test_feature_detection_is_refused_on_a_row_band():
    gv = A.Video(32, 16, 1)
    gv.set_row_offset(16)
    gv.update_detect_features(True)
    with pytest.raises(A.AdderError) as e:
        gv.integrate_matrix(np.zeros((16, 32, 1), np.uint8), 255.0)
    assert e.value.code == B.ERR_UNSUPPORTED


def test_host_form_resumes_after_a_full_output_buffer():
    """events_cap too small in the middle of a batch.  The call reports the frames it delivered; the frames the
    pipeline had already integrated are neither lost nor integrated twice: resuming from frames_done yields the oracle's
    stream for every frame, and the state agrees at the end."""
    case = cases.CASES_BY_NAME["cfg2_rgb_noise_crf3"]
    gv, ov = _pair(case)
    frames = case.frames()
    exp = [ov.integrate_matrix(frames[f], case.time)[0] for f in range(case.n_frames)]
    sizes = [len(e) for e in exp]
    small = np.empty(sum(sizes[:7]) + sizes[7] // 2, dtype=A.EVENT_DTYPE)  # room for 7 frames and a half
    got, f0, calls = [], 0, 0
    while f0 < case.n_frames:
        ev, fc, cc, done = gv.integrate_frames_host(frames[f0:], case.time, small, partial=True)
        calls += 1
        assert done >= 1, "no progress"
        assert [int(x) for x in fc] == sizes[f0:f0 + done]
        got.append(ev.copy())
        f0 += done
        if calls == 1:  # the kept frames block the other entry points until the call is resumed
            assert done == 7
            with pytest.raises(A.AdderError) as e:
                gv.integrate_matrix(frames[0], case.time)
            assert e.value.code == A.binding.ERR_BAD_PARAMS
    assert calls >= 3
    assert np.concatenate(got).tobytes() == np.concatenate(exp).tobytes()
    n = case.w * case.h * case.c
    _assert_state_equal(gv, ov, n, step=3)
    assert gv.info().in_interval_count == ov.in_interval_count


def test_normal_mode_long_static_run_needs_no_depth_hint():
    """PixelMultiMode::Normal never pops the root again after the first Δt_max pop, so a static pixel's stack
    keeps growing past what delta_t_max / ref_time suggests (11-13 live nodes after 5000 frames at dtm = ref).  The
    handle derives the depth from the mode: no ADDER_ERR_ARENA_DEPTH, and the oracle's events."""
    w, h, nf, batch = 24, 8, 5000, 500
    rng = np.random.default_rng(7)
    base = rng.integers(0, 256, (h, w, 1)).astype(np.uint8)
    gv = A.Video(w, h, 1)
    ov = O.Video(w, h, 1, O.MODE_FRAME_PERFECT)
    for v in (gv, ov):
        assert v.time_parameters(255 * 30, 255, 255, None)  # dtm = ref: the shallowest derived depth
        v.write_out(None, O.MULTI_NORMAL)
    frames = np.broadcast_to(base[None], (batch, h, w, 1)).copy()
    buf = np.empty(w * h * 4 * batch, dtype=A.EVENT_DTYPE)
    for f0 in range(0, nf, batch):
        ev, fc, cc = gv.integrate_frames_host(frames, 255.0, buf)
        exp = [ov.integrate_matrix(base, 255.0)[0] for _ in range(batch)]
        assert [int(x) for x in fc] == [len(e) for e in exp]
        assert ev.tobytes() == np.concatenate(exp).tobytes()
    deepest = max(ov.px(i).length for i in range(w * h))
    assert deepest >= 11, deepest
    _assert_state_equal(gv, ov, w * h)
