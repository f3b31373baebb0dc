"""A non-Python consumer of the C ABI: tests/c_abi/consumer.c is plain C11 compiled against include/adder_b200.h and
linked with libadder_b200.so.  CPU part: it compiles without warnings as C (so the header is C, not C++), its static
layout assertions hold, and without a GPU the library refuses loudly.  GPU part: it runs BASELINE configs[0] (640x480
gray gradient, 30 frames, Video::new defaults), builds the reference's per-chunk vectors, and what it read back from those
vectors is the oracle's stream."""
import json
import os
import subprocess

import numpy as np
import pytest

import adder_codec_rs_b200 as A
from oracle import oracle_py as O
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "consumer.c")


def _build(tmp_path):
    A.build()
    exe = str(tmp_path / "consumer")
    libdir = os.path.join(ROOT, "adder_codec_rs_b200")
    subprocess.run(["gcc", "-std=c11", "-O2", "-pthread", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), SRC,
                    "-L", libdir, "-ladder_b200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def test_header_is_plain_c_and_the_consumer_links(tmp_path):
    exe = _build(tmp_path)
    if A.device_count() == 0:  # no GPU here: the product must fail loudly, not fall back
        r = subprocess.run([exe, "cfg1", "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU path" in r.stderr


def _read_out(path, n_frames):
    raw = open(path, "rb").read()
    pos, frames = 0, []
    for _ in range(n_frames):
        n_chunks = int(np.frombuffer(raw, np.uint32, 1, pos)[0])
        pos += 4
        counts = np.frombuffer(raw, np.uint32, n_chunks, pos)
        pos += 4 * n_chunks
        n = int(counts.sum())
        ev = np.frombuffer(raw, A.EVENT_DTYPE, n, pos)
        pos += 12 * n
        frames.append((counts, ev))
    assert pos == len(raw)
    return frames


@pytest.mark.gpu
def test_c_consumer_runs_cfg1_and_matches_the_oracle(tmp_path):
    exe = _build(tmp_path)
    out = str(tmp_path / "cfg1.bin")
    r = subprocess.run([exe, "cfg1", "30", out, "pinned", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    stats = json.loads(r.stdout.strip().splitlines()[-1])
    ov = O.Video(640, 480, 1, O.MODE_FRAME_PERFECT)
    assert ov.time_parameters(255 * 30, 255, 255, None)
    total = 0
    for f, (counts, ev) in enumerate(_read_out(out, 30)):
        eo, co = ov.integrate_matrix(synth.frame(synth.GRADIENT, 0xADDE5, f, 640, 480, 1), 255.0)
        assert np.array_equal(counts, co), f"frame {f}: per-chunk vector lengths"
        assert ev.tobytes() == eo.tobytes(), f"frame {f}: events read back from the per-chunk vectors"
        total += len(eo)
    assert stats["events"] == total and total > 0
    print("C consumer, cfg1:", stats)


@pytest.mark.gpu
def test_c_consumer_cost_of_one_consume_at_1080p_rgb(tmp_path):
    """The number VERDICT r1 asked for: one consume() at 1080p RGB with the Vec<Vec<Event>> build included (first frames
    compared with the oracle as well)."""
    exe = _build(tmp_path)
    out = str(tmp_path / "cfg2.bin")
    r = subprocess.run([exe, "cfg2", "12", out, "pinned", str(len(os.sched_getaffinity(0)))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    stats = json.loads(r.stdout.strip().splitlines()[-1])
    ov = O.Video(1920, 1080, 3, O.MODE_FRAME_PERFECT)
    assert ov.time_parameters(255 * 30, 255, 7650, None)
    ov.update_crf(3)
    for f, (counts, ev) in enumerate(_read_out(out, 12)[:3]):
        eo, co = ov.integrate_matrix(O.synth_frame(synth.NOISE, 0xADDE5, f, 1920, 1080, 3), 255.0, O.host_threads())
        assert np.array_equal(counts, co) and ev.tobytes() == eo.tobytes(), f"frame {f}"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "c_consumer_1080p.json"), "w") as fh:
        json.dump(stats, fh)
    print("C consumer, cfg2:", stats)
