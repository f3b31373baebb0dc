#!/usr/bin/env python
"""Small transcode runs for compute-sanitizer (memcheck / racecheck / synccheck): a few shared cases through all
three forms of the step, checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adder_codec_rs_b200 as A  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import cases  # noqa: E402

for name in sys.argv[1:] or ["cfg2_rgb_noise_crf3", "jitter_dtm4_normal", "ragged_37x13x3_chunk4"]:
    case = cases.CASES_BY_NAME[name]
    gv = A.Video(case.w, case.h, case.c)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(gv, case)
    cases.configure(ov, case)
    frames = case.frames()
    n = min(case.n_frames - 4, 8)
    for f in range(n):
        eg, cg = gv.integrate_matrix(frames[f], case.time)
        eo, co = ov.integrate_matrix(frames[f], case.time)
        assert eg.tobytes() == eo.tobytes() and np.array_equal(cg, co), (name, f)
    out = np.empty(1 << 20, dtype=np.uint8)
    body, _, _ = gv.integrate_frames_host_raw(np.ascontiguousarray(frames[n:n + 4]), case.time, out)
    want = b"".join(O.raw_encode(ov.integrate_matrix(frames[f], case.time)[0], case.c) for f in range(n, n + 4))
    assert body.tobytes() == want, name
    print("ok", name, flush=True)
