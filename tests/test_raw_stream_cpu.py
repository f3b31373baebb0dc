"""The raw .adder wire format restated in the oracle against the
reference's own fixture: re-serialising the golden events must reproduce the whole file."""
import hashlib
import json
import os
import struct

import numpy as np

from oracle import oracle_py as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_events():
    g = np.load(os.path.join(G, "lake_events.npz"))
    ev = np.zeros(len(g["x"]), dtype=O.EVENT_DTYPE)
    ev["x"], ev["y"], ev["d"], ev["t"] = g["x"], g["y"], g["d"], g["t"]
    ev["c"] = O.C_NONE
    return ev, g["header"].tobytes()


def test_reserialised_fixture_is_byte_identical():
    ev, header = _golden_events()
    dig = json.load(open(os.path.join(G, "lake_adder_digest.json")))
    # header fields as the fixture declares them (Appendix C): 200x50x1, tps 6113, ref 255, dtm 6120, v3, FramedU8, DeltaT
    w, h, tps, ref, dtm = struct.unpack(">HHIII", header[7:23])
    source, time_mode, adu = struct.unpack(">III", header[25:37])
    mine = O.raw_header(w, h, header[24], tps, ref, dtm, version=header[5], source_camera=source, time_mode=time_mode, adu_interval=adu)
    assert mine == header
    blob = mine + O.raw_encode(ev, 1) + O.raw_eof()
    assert len(blob) == dig["size"] == 37 + 9 * len(ev) + 11
    assert hashlib.sha256(blob).hexdigest() == dig["sha256"]


def test_header_sizes_per_version():
    """tests/integration_tests.rs:216-245 pins 29 (v1) and 33 (v2); v3 is 37 (fixture), v0 25."""
    sizes = [len(O.raw_header(64, 48, 1, 7650, 255, 7650, version=v)) for v in range(4)]
    assert sizes == [25, 29, 33, 37]
    assert O.raw_header(1, 1, 1, 1, 1, 1, version=4) == b""
    h = O.raw_header(640, 480, 3, 7650, 255, 7650, version=2, time_mode=O.TIME_ABSOLUTE_T)
    assert h[:5] == b"adder" and h[23] == 11 and h[24] == 3 and h[29:33] == b"\x00\x00\x00\x01"
    assert O.raw_header(640, 480, 3, 7650, 255, 7650, compressed=True)[:5] == b"addec"


def test_colour_events_are_eleven_bytes_with_option_tag():
    ev = np.zeros(2, dtype=O.EVENT_DTYPE)
    ev[0] = (0x0102, 0x0304, 2, 7, 0, 0x0A0B0C0D)
    ev[1] = (65534, 1, 0, 255, 0, 0xFFFFFFFF)
    b = O.raw_encode(ev, 3)
    assert b[:11] == bytes([1, 2, 3, 4, 1, 2, 7, 0x0A, 0x0B, 0x0C, 0x0D])
    assert b[11:] == bytes([0xFF, 0xFE, 0, 1, 1, 0, 255, 0xFF, 0xFF, 0xFF, 0xFF])
    assert O.raw_encode(ev, 1)[:9] == bytes([1, 2, 3, 4, 7, 0x0A, 0x0B, 0x0C, 0x0D])
    assert O.raw_eof() == bytes([0xFF, 0xFF, 0xFF, 0xFF, 1, 0, 0, 0, 0, 0, 0])


def test_handle_color_matches_the_formula():
    """oracle_handle_color == (ch0*0.114 + ch1*0.587 + ch2*0.299) as u8 in f64 (utils/cv.rs:215-232), all 2^24 triples sampled."""
    def handle_color(frame):  # the formula in numpy f64 (a test-side statement: the product converts on the device)
        f = frame.astype(np.float64)
        g = f[..., 0] * 0.114 + f[..., 1] * 0.587 + f[..., 2] * 0.299
        return np.clip(np.trunc(g), 0, 255).astype(np.uint8)[..., None]

    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (64, 257, 3), dtype=np.uint8)
    rgb[0, :8] = [[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 255], [128, 128, 128]]
    want = handle_color(rgb)
    assert np.array_equal(O.handle_color(rgb), want)
    assert want[0, 0, 0] == 255 and want[0, 1, 0] == 0  # 0.114 + 0.587 + 0.299 = 1.0 exactly at 255? (254.99999999999997 -> 254 would show here)
