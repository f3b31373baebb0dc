"""The device framer (INSTANTANEOUS, u8) through the C ABI: the reference's golden pairs straight through the GPU,
and lock-step transcoder -> framer runs against the oracle pair (GPU box only)."""
import os

import numpy as np
import pytest

import adder_codec_rs_b200 as A
from oracle import oracle_py as O
from tests import cases, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["ordered", "unordered"])
def test_sample_3_golden_frames_through_the_gpu(name):
    """tests/integration_tests.rs:818-962: one event at a time, write_multi_frame_bytes whenever a frame is ready -> 405 frames == sample_3.gray"""
    g = np.load(os.path.join(G, "framer_sample3.npz"))
    fr = A.Framer(10, 5, 1, 64, 0, A.TIME_DELTA_T, 300000, 5000, 3000000, output_fps=60.0, ring_frames=700)
    assert fr.tpf == 5000
    frames = []
    x, y, d, t = (g[f"{name}_{k}"] for k in "xydt")
    for i in range(len(x)):
        if fr.ingest_event(int(x[i]), int(y[i]), 0xFF, int(d[i]), int(t[i])):
            got = fr.write_multi_frame_bytes()
            assert len(got) > 0, "should have frame"
            frames.append(got)
    frames = np.concatenate(frames)
    assert len(frames) == 405
    assert np.array_equal(frames, g["gray"])


def test_lake_golden_frames_through_the_gpu():
    """adder_simulproc.rs:169-268 `dark`: the lake events, one transcoded frame per ingest_events_events -> lake_scaled_out."""
    g = np.load(os.path.join(G, "lake_events.npz"))
    want = np.load(os.path.join(G, "lake_scaled_out.npy"))
    w, h = 200, 50
    ev = np.zeros(len(g["x"]), dtype=A.EVENT_DTYPE)
    ev["x"], ev["y"], ev["d"], ev["t"], ev["c"] = g["x"], g["y"], g["d"], g["t"], 0xFF
    fps = float(np.float32(24000.0 / 1001.0))
    fr = A.Framer(w, h, 1, 1, 3, A.TIME_DELTA_T, 6113, 255, 6120, output_fps=fps, ring_frames=160)  # zero-intensity pixels lag ~100 frames in this stream
    of = O.Framer(w, h, 1, 1, 3, O.TIME_DELTA_T, 6113, 255, 6120, output_fps=fps)
    key = ev["y"].astype(np.int64) * w + ev["x"]
    bounds = np.concatenate([[0], np.flatnonzero(np.diff(key) < 0) + 1, [len(key)]])
    frames = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        part = ev[a:b]
        counts = np.bincount(part["y"], minlength=h).astype(np.uint32)
        ready = fr.ingest_events_events(part, counts)
        assert ready == of.ingest_events_events(part, counts)
        if ready:
            got = fr.write_multi_frame_bytes()
            assert np.array_equal(got, of.write_multi_frame_bytes())
            frames.append(got)
    frames = np.concatenate(frames)
    assert np.array_equal(frames[: len(want)], want)


LOCKSTEP = [
    # name, w, h, c, kind, frames, case kwargs, framer kwargs
    ("noise_rgb_abs", 40, 24, 3, synth.NOISE, 50, dict(crf=3), dict()),
    ("jitter_deltat_normal", 48, 16, 1, synth.JITTER, 80, dict(manual=(12, 12, 4, 1), dtm=255 * 4, time_mode=O.TIME_DELTA_T, multi_mode=O.MULTI_NORMAL), dict()),
    ("static_chunk4_limit", 37, 13, 1, synth.STATIC_BLIPS, 90, dict(crf=5, dtm=255 * 16, chunk_rows=4), dict(buffer_limit=6)),
    ("gradient_view_dt", 64, 16, 1, synth.GRADIENT, 60, dict(manual=(40, 40, 30, 1)), dict(view_mode=O.VIEW_DELTA_T)),
    ("noise_30fps_out_of_60", 32, 16, 1, synth.NOISE, 60, dict(crf=3), dict(fps_scale=0.5)),
]


@pytest.mark.parametrize("spec", LOCKSTEP, ids=lambda s: s[0])
def test_transcoder_and_framer_in_lock_step(spec):
    """SimulProcessor's loop (simulproc.rs:229-277): every consume() goes to ingest_events_events, frames are written
    when ready, the buffer is flushed at the end — device transcoder + device framer against oracle + oracle."""
    name, w, h, c, kind, nf, ckw, fkw = spec
    fkw = dict(fkw)
    case = cases.Case(name, w, h, c, kind, nf, **ckw)
    gv = A.Video(w, h, c)
    ov = O.Video(w, h, c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    time_mode = O.TIME_ABSOLUTE_T if case.time_mode is None else case.time_mode
    source_fps = 30.0
    tps = case.ref * 30
    out_fps = source_fps * fkw.pop("fps_scale", 1.0)
    args = (w, h, c, case.chunk_rows, 3, time_mode, tps, case.ref, case.dtm)
    gf = A.Framer(*args, output_fps=out_fps, ring_frames=400, **fkw)  # pixels that sit at intensity 0 hold the front frame back
    of = O.Framer(*args, output_fps=out_fps, **fkw)
    frames = case.frames()
    P = w * h * c
    d_frame = gv.device_alloc(P)
    cap = P * 4
    d_events = gv.device_alloc(cap * 12)
    d_off = gv.device_alloc((gv.n_chunks + 1) * 4)
    n_out = 0
    for f in range(nf):
        d_frame.from_host(frames[f])
        gv.integrate_frames_device(d_frame.ptr, P, 1, case.time, d_events.ptr, cap, d_off.ptr)
        gv.sync()
        eo, co = ov.integrate_matrix(frames[f], case.time)
        if f % 2:  # the asynchronous form: kernels queued, the predicate fetched by a separate call
            gf.ingest_events_device_async(d_events.ptr, d_off.ptr)
            ready_g = gf.frame_ready()
        else:
            ready_g = gf.ingest_events_device(d_events.ptr, d_off.ptr)  # events never leave HBM
        ready_o = of.ingest_events_events(eo, co)
        assert ready_g == ready_o, f"frame {f}"
        if ready_o:
            a, b = gf.write_multi_frame_bytes(), of.write_multi_frame_bytes()
            assert np.array_equal(a, b), f"frame {f}: reconstructed frames differ"
            n_out += len(b)
    assert gf.flush_frame_buffer() == of.flush_frame_buffer()
    a, b = gf.write_multi_frame_bytes(), of.write_multi_frame_bytes()
    assert np.array_equal(a, b)
    n_out += len(b)
    assert gf.frames_written == of.frames_written == n_out
    assert n_out >= 1 and not of.bad


def test_ring_overflow_is_reported():
    fr = A.Framer(8, 4, 1, 64, 1, A.TIME_DELTA_T, 50000, 1000, 1000, output_fps=50.0, ring_frames=4)
    with pytest.raises(A.AdderError) as e:
        fr.ingest_event(1, 1, 0xFF, 5, 50000)  # reaches 50 frames ahead
    assert e.value.code == A.binding.ERR_CAPACITY


def test_simulproc_mirror_writes_the_oracle_pipeline_frames_and_stream(tmp_path):
    """Framed -> SimulProcessor.run: reconstructed frames file and raw .adder file against the oracle transcoder + framer
    + raw encoder driven the same way (colour source, gray transcode, so handle_color is in the loop too)."""
    import io

    w, h, nf = 48, 20, 45
    rgb = synth.moving_blocks(9, nf, w, h, 3)
    src = A.Framed(list(rgb), w, h, color_input=False, source_fps=24.0)
    src = src.crf(2).auto_time_parameters(255, 255 * 8, None).write_out(A.TIME_ABSOLUTE_T, A.MULTI_COLLAPSE)
    frames_io, raw_io = io.BytesIO(), io.BytesIO()
    from tools.simulproc_mirror import SimulProcessor

    sp = SimulProcessor(src, 255, frames_io, raw_output=raw_io, ring_frames=200)
    assert sp.run() == nf
    # the oracle, step by step
    ov = O.Video(w, h, 1, O.MODE_FRAME_PERFECT)
    ov.update_crf(2)
    assert ov.time_parameters(int(255 * 24), 255, 255 * 8, None)
    ov.write_out(O.TIME_ABSOLUTE_T, O.MULTI_COLLAPSE)
    of = O.Framer(w, h, 1, 1, 3, O.TIME_ABSOLUTE_T, int(255 * 24), 255, 255 * 8, output_fps=24.0)
    want_frames, want_raw = b"", O.raw_header(w, h, 1, int(255 * 24), 255, 255 * 8, version=3, time_mode=O.TIME_ABSOLUTE_T)
    for f in range(nf):
        eo, co = ov.integrate_matrix(O.handle_color(rgb[f]), 255.0)
        want_raw += O.raw_encode(eo, 1)
        if of.ingest_events_events(eo, co):
            want_frames += of.write_multi_frame_bytes().tobytes()
    want_raw += O.raw_eof()
    assert raw_io.getvalue() == want_raw
    assert frames_io.getvalue() == want_frames and len(want_frames) >= w * h * 3
