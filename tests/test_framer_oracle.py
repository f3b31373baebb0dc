"""The framer oracle (oracle/framer_oracle.c) against the reference's own golden pairs (CPU only):
 - tests/integration_tests.rs:818-962 test_sample_{ordered,unordered}: sample_3_*.adder -> sample_3.gray, 405 frames,
   fed one event at a time through ingest_event + write_multi_frame_bytes exactly like the reference test;
 - adder_simulproc.rs:169-268 `dark`: the lake .adder events -> lake_scaled_out (11 frames), fed per transcoded
   frame through ingest_events_events like SimulProcessor (simulproc.rs:166-218)."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["ordered", "unordered"])
def test_sample_3_reconstructs_the_golden_frames(name):
    g = np.load(os.path.join(G, "framer_sample3.npz"))
    # header of the fixture: v0, 10x5x1, tps 300000, ref 5000, dtm 3000000; the test frames at 60 fps, chunk_rows 64, DeltaT
    fr = O.Framer(10, 5, 1, 64, 0, O.TIME_DELTA_T, 300000, 5000, 3000000, output_fps=60.0)
    assert fr.tpf == 5000
    frames = []
    x, y, d, t = (g[f"{name}_{k}"] for k in "xydt")
    for i in range(len(x)):
        if fr.ingest_event(int(x[i]), int(y[i]), O.C_NONE, int(d[i]), int(t[i])):
            got = fr.write_multi_frame_bytes()
            assert len(got) > 0, "should have frame"
            frames.append(got)
    assert not fr.bad
    frames = np.concatenate(frames)
    assert len(frames) == 405  # assert_eq!(frame_count, 405)
    assert np.array_equal(frames, g["gray"])


def _split_frames(key):
    """Indices where the raster key restarts: the per-frame groups SimulProcessor handed to the framer."""
    cuts = np.flatnonzero(np.diff(key.astype(np.int64)) < 0) + 1
    return np.concatenate([[0], cuts, [len(key)]])


def test_lake_events_reconstruct_lake_scaled_out():
    g = np.load(os.path.join(G, "lake_events.npz"))
    want = np.load(os.path.join(G, "lake_scaled_out.npy"))
    w, h = 200, 50
    ev = np.zeros(len(g["x"]), dtype=O.EVENT_DTYPE)
    ev["x"], ev["y"], ev["d"], ev["t"], ev["c"] = g["x"], g["y"], g["d"], g["t"], O.C_NONE
    # SimulProcessor::new (simulproc.rs:144-161): codec_version 3 (the fixture's header), DeltaT, tps 6113, ref 255,
    # dtm 6120, output fps = the source's 23.976 fps, chunk_rows 1 (VideoState::default)
    fr = O.Framer(w, h, 1, 1, 3, O.TIME_DELTA_T, 6113, 255, 6120, output_fps=float(np.float32(24000.0 / 1001.0)))
    key = ev["y"].astype(np.int64) * w + ev["x"]
    bounds = _split_frames(key)
    frames = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        part = ev[a:b]
        counts = np.bincount(part["y"], minlength=h).astype(np.uint32)
        if fr.ingest_events_events(part, counts):
            got = fr.write_multi_frame_bytes()
            assert len(got) > 0
            frames.append(got)
    assert not fr.bad
    frames = np.concatenate(frames)
    print("frames reconstructed", len(frames), "tpf", fr.tpf)
    assert len(frames) >= len(want)
    assert np.array_equal(frames[: len(want)], want)


def test_fill_logic_of_the_reference_unit_tests():
    """tests/integration_tests.rs:459-495 test_event_framer_ingest_get_filled and the doc example driver.rs:404-436."""
    fr = O.Framer(5, 5, 1, 64, 1, O.TIME_DELTA_T, 50000, 1000, 1000, output_fps=50.0)
    for i in range(5):
        for j in range(5):
            filled = fr.ingest_event(i, j, O.C_NONE, 5, 5100)
            assert filled == (i == 4 and j == 4)
    fr2 = O.Framer(10, 10, 3, 64, 1, O.TIME_DELTA_T, 50000, 1000, 1000, output_fps=50.0)
    fr2.ingest_event(5, 5, 1, 5, 1000)
    fr2.flush_frame_buffer()
    # px_at_current(5, 5, 1) == Some(32): 2^5 / 1000 * 1000
    for k in range(10):
        for j in range(10):
            for c in range(3):
                if (k, j, c) != (5, 5, 1):
                    fr2.ingest_event(j, k, c, 0, 1000)
    got = fr2.write_multi_frame_bytes()
    assert got[0, 5, 5, 1] == 32
