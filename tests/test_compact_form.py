"""The compact host form (include/adder_b200.h): per frame either P count bytes + E x {d, t} or E x {index, d, t}.
CPU: the host-side expander against blocks built in numpy from oracle streams.  GPU: integrate_frames_host_compact ->
expand_compact == the oracle's records, for dense (noise), sparse (static) and banded planes."""
import numpy as np
import pytest

import adder_codec_rs_b200 as A
from oracle import oracle_py as O
from tests import cases


def _block_of(ev, w, h, c, row0=0):
    """numpy statement of the compact form for one frame's records."""
    P = w * h * c
    idx = ((ev["y"].astype(np.int64) - row0) * w + ev["x"]) * c + (0 if c == 1 else ev["c"].astype(np.int64))
    dt = np.zeros(len(ev), dtype=np.dtype([("d", "u1"), ("t", "<u4")], align=False))
    dt["d"], dt["t"] = ev["d"], ev["t"]
    if 4 * len(ev) > P:
        counts = np.bincount(idx, minlength=P).astype(np.uint8)
        return np.concatenate([counts, dt.view(np.uint8)])
    rec = np.zeros(len(ev), dtype=np.dtype([("i", "<u4"), ("d", "u1"), ("t", "<u4")], align=False))
    rec["i"], rec["d"], rec["t"] = idx, ev["d"], ev["t"]
    return rec.view(np.uint8)


@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "cfg5_static_collapse", "cfg3_jitter_c10", "ragged_37x13x3_chunk4"])
def test_expander_inverts_the_numpy_statement_of_the_form(name):
    case = cases.CASES_BY_NAME[name]
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    frames = case.frames()
    forms = set()
    for f in range(min(case.n_frames, 40)):
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        block = _block_of(eo, case.w, case.h, case.c)
        assert len(block) == A.compact_frame_bytes(case.w * case.h * case.c, len(eo))
        forms.add(4 * len(eo) > case.w * case.h * case.c)
        for nt in (1, 3, 8):
            got = A.expand_compact(case.w, case.h, case.c, 0, block, len(eo), n_threads=nt)
            assert got.tobytes() == eo.tobytes(), f"frame {f}, {nt} threads"
    assert forms, "no frames"


def test_expander_rejects_a_block_that_does_not_add_up():
    block = np.zeros(16 + 5 * 8, dtype=np.uint8)  # dense form for P = 16, E = 8, but all counts are zero
    with pytest.raises(A.AdderError):
        A.expand_compact(4, 4, 1, 0, block, 8)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "cfg5_static_collapse", "cfg3_jitter_c5", "ragged_37x13x3_chunk4", "jitter_dtm4_normal"])
def test_compact_host_form_expands_to_the_oracle_stream(name):
    case = cases.CASES_BY_NAME[name]
    gv = A.Video(case.w, case.h, case.c)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    frames = case.frames()
    P = case.w * case.h * case.c
    buf = np.empty(P * 12 * case.n_frames, dtype=np.uint8)
    body, fc, cc = gv.integrate_frames_host_compact(frames, case.time, buf)
    pos = 0
    dense = sparse = 0
    for f in range(case.n_frames):
        eo, co = ov.integrate_matrix(frames[f], case.time)
        assert int(fc[f]) == len(eo) and np.array_equal(cc[f], co), f"frame {f}"
        nb = A.compact_frame_bytes(P, len(eo))
        got = A.expand_compact(case.w, case.h, case.c, 0, body[pos:pos + nb], len(eo), n_threads=4)
        assert got.tobytes() == eo.tobytes(), f"frame {f}"
        assert body[pos:pos + nb].tobytes() == _block_of(eo, case.w, case.h, case.c).tobytes(), f"frame {f}: block bytes"
        pos += nb
        dense += 4 * len(eo) > P
        sparse += 4 * len(eo) <= P
    assert pos == len(body)
    print(name, "dense frames", dense, "sparse frames", sparse)


@pytest.mark.gpu
def test_compact_form_of_a_row_band_carries_band_local_indices():
    from adder_codec_rs_b200 import sharding as S

    case = cases.CASES_BY_NAME["cfg2_rgb_noise_crf3"]
    bands = [S.BandedVideo(case.w, case.h, case.c, r, 2, device=0) for r in range(2)]
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    for b in bands:
        cases.configure(b, case)
    frames = case.frames()[:12]
    got = [[] for _ in range(len(frames))]
    for b in bands:
        P = case.w * b.rows * case.c
        buf = np.empty(P * 12 * len(frames), dtype=np.uint8)
        body, fc, cc = b.integrate_frames_host_compact(np.ascontiguousarray(frames[:, b.row0:b.row0 + b.rows]), case.time, buf)
        pos = 0
        for f in range(len(frames)):
            nb = A.compact_frame_bytes(P, int(fc[f]))
            got[f].append(A.expand_compact(case.w, b.rows, case.c, b.row0, body[pos:pos + nb], int(fc[f])))
            pos += nb
    for f in range(len(frames)):
        eo, _ = ov.integrate_matrix(frames[f], case.time)
        assert np.concatenate(got[f]).tobytes() == eo.tobytes(), f"frame {f}"
