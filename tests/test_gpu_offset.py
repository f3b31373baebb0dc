"""The offset form of the node stacks on the GPU (csrc/px_offset.cuh behind integrate_frame_kernel<.., kOff>): which
launches use it, long runs against the oracle, and the conversion back to the reference's representation when a launch
stops being eligible.  GPU box only; the same header runs on the host in tests/test_px_offset_host.py."""
import numpy as np
import pytest

import adder_codec_rs_b200 as A
from oracle import oracle_py as O
from tests import cases, synth
from tests.cases import Case
from tests.test_gpu_parity import _assert_state_equal, _pair
from tests.test_px_offset_host import LONG, _random_frames

pytestmark = pytest.mark.gpu


def _step(gv, ov, frame, time, f):
    eg, cg = gv.integrate_matrix(frame, time)
    eo, co = ov.integrate_matrix(frame, time)
    assert len(eg) == len(eo), f"frame {f}: {len(eg)} vs {len(eo)} events"
    assert eg.tobytes() == eo.tobytes(), f"frame {f}: event streams differ"
    assert np.array_equal(cg, co), f"frame {f}: chunk counts differ"
    return len(eg)


def test_which_launches_use_the_offset_form():
    served = {}
    for case in cases.CASES:
        gv, _ = _pair(case)
        gv.integrate_matrix(case.frames(0, 1)[0], case.time)
        served[case.name] = gv.state_form
    assert served["cfg5_static_collapse"] == 1 and served["cfg2_rgb_noise_crf3"] == 1 and served["cfg3_jitter_c10"] == 1
    assert served["cfg1_gradient_dtm_eq_ref"] == 1 and served["initial_d"] == 1 and served["roi_rect"] == 1
    assert served["cfg5_static_normal"] == 0 and served["time_spanned_fraction"] == 0 and served["dark_params_jitter"] == 0


@pytest.mark.parametrize("case", LONG, ids=lambda c: c.name)
def test_long_runs_in_offset_form(case):
    """Frame by frame (single-frame launches), state compared every 97 frames and at the end."""
    gv, ov = _pair(case)
    n = case.w * case.h * case.c
    total = 0
    for f0 in range(0, case.n_frames, 100):
        frames = case.frames(f0, min(100, case.n_frames - f0))
        for k in range(len(frames)):
            total += _step(gv, ov, frames[k], case.time, f0 + k)
            if (f0 + k) % 97 == 0:
                _assert_state_equal(gv, ov, n)
    assert gv.state_form == 1
    assert total > 0
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    _assert_state_equal(gv, ov, n)


@pytest.mark.parametrize("seed", range(6))
def test_random_scenes_through_multi_frame_launches(seed):
    """Random parameters and scenes (tests/test_px_offset_host.py) through integrate_frames_device in launches of 1..40 frames."""
    rng = np.random.default_rng(4000 + seed)
    ref = int(rng.choice([255, 256, 1000]))
    mult = int(rng.choice([2, 5, 13, 30, 200]))
    c_base = int(rng.choice([0, 4, 10, 25]))
    case = Case(f"rand{seed}", 40, 9, 1 if seed % 2 else 3, 0, 300, manual=(c_base, c_base + int(rng.choice([0, 20])), mult, int(rng.choice([1, 3]))),
                ref=ref, dtm=ref * mult)
    frames = _random_frames(rng, case.n_frames, case.h, case.w, case.c)
    gv, ov = _pair(case)
    P = case.w * case.h * case.c
    cap = P * 8
    d_frames = gv.device_alloc(P * 40)
    d_events = gv.device_alloc(cap * 12 * 40)
    nck = gv.n_chunks
    d_off = gv.device_alloc((nck + 1) * 4 * 40)
    f = 0
    while f < case.n_frames:
        n = min(int(rng.integers(1, 41)), case.n_frames - f)
        d_frames.from_host(frames[f:f + n])
        gv.integrate_frames_device(d_frames.ptr, P, n, case.time, d_events.ptr, cap, d_off.ptr)
        gv.sync()
        offs = d_off.to_host(np.uint32, nbytes=(nck + 1) * 4 * n).reshape(n, nck + 1)
        for k in range(n):
            eo, co = ov.integrate_matrix(frames[f + k], case.time)
            assert offs[k, -1] == len(eo), f"frame {f + k}: {offs[k, -1]} vs {len(eo)} events"
            assert np.array_equal(np.diff(offs[k]), co)
            eg = d_events.to_host(A.EVENT_DTYPE, nbytes=len(eo) * 12, offset=k * cap * 12)
            assert eg.tobytes() == eo.tobytes(), f"frame {f + k}: event streams differ"
        f += n
    for b in (d_frames, d_events, d_off):
        b.free()
    assert gv.state_form == 1
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    _assert_state_equal(gv, ov, P)


@pytest.mark.parametrize("how", ["normal_mode", "fractional_time", "huge_dtm"])
def test_state_is_converted_when_a_launch_stops_being_eligible(how):
    """Deep offset-form stacks, then a parameter change the form cannot serve: the library rewrites the state in the
    reference's representation (offset_to_eager_kernel) and carries on with px_step; the oracle sees the same calls."""
    case = Case("convert", 24, 10, 1, synth.STATIC_BLIPS, 260, crf=3, ref=256, dtm=256 * 300)
    gv, ov = _pair(case)
    frames = case.frames()
    n = case.w * case.h * case.c
    time = case.time
    for f in range(case.n_frames):
        if f == 130:
            assert gv.state_form == 1
            assert max(gv.px_dict(i)["length"] for i in range(n)) >= 5
            if how == "normal_mode":
                gv.write_out(None, A.MULTI_NORMAL)
                ov.write_out(None, O.MULTI_NORMAL)
            elif how == "fractional_time":
                time = 100.5
            else:
                gv.update_delta_t_max(256 * 40000)
                ov.update_delta_t_max(256 * 40000)
        _step(gv, ov, frames[f], time, f)
        if f == 130:
            assert gv.state_form == 0
            _assert_state_equal(gv, ov, n)
    assert gv.state_form == 0  # an eager state that has integrated frames stays eager
    assert np.array_equal(gv.running_intensities(), ov.running_intensities())
    _assert_state_equal(gv, ov, n)


def test_reset_state_lets_the_form_be_chosen_again():
    case = cases.CASES_BY_NAME["cfg5_static_collapse"]
    gv, ov = _pair(case)
    frames = case.frames(0, 12)
    gv.write_out(None, A.MULTI_NORMAL)
    gv.integrate_matrix(frames[0], case.time)
    assert gv.state_form == 0
    gv.write_out(None, A.MULTI_COLLAPSE)
    gv.integrate_matrix(frames[1], case.time)
    assert gv.state_form == 0
    gv.reset_state()
    gv.update_crf(case.crf)  # reset_state gives Video::new's c_thresh back
    for f in range(12):
        _step(gv, ov, frames[f], case.time, f)
    assert gv.state_form == 1
    _assert_state_equal(gv, ov, case.w * case.h * case.c)
