"""The offset form of the node stacks (adder_codec_rs_b200/csrc/px_offset.cuh) on the host simulation against the oracle:
long runs that reach deep stacks, Δt_max pops and the frozen regime, with the state converted back to the reference's
nodes and compared too.  CPU only; the GPU runs the same header (tests/test_gpu_offset.py)."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import cases, sim_py, synth
from tests.cases import Case
from tests.test_px_machine_host import _SimAdapter

LONG = [
    # BASELINE config 5: static scene with blips, ref 256, dtm 2^20 -- run PAST the first Δt_max pop (frame 4096)
    Case("off_cfg5_static_4300", 12, 6, 1, synth.STATIC_BLIPS, 4300, crf=3, ref=256, dtm=1 << 20),
    # the same scene with a short dtm: many pops, shifts and frozen stretches
    Case("off_static_dtm64", 16, 8, 1, synth.STATIC_BLIPS, 700, crf=3, ref=256, dtm=256 * 64),
    Case("off_static_dtm7_rgb", 8, 8, 3, synth.STATIC_BLIPS, 400, crf=6, ref=255, dtm=255 * 7),
    # config 3: jitter at every c of the sweep's upper half and beyond (unchanged pixels for long stretches)
    Case("off_jitter_c7", 24, 12, 1, synth.JITTER, 200, manual=(7, 7, 30, 1)),
    Case("off_jitter_c10", 24, 12, 1, synth.JITTER, 200, manual=(10, 10, 30, 1)),
    Case("off_jitter_c15", 24, 12, 1, synth.JITTER, 300, manual=(15, 15, 30, 1)),
    Case("off_jitter_c20_dtm200", 24, 12, 1, synth.JITTER, 500, manual=(20, 20, 200, 1), dtm=255 * 200),
    Case("off_jitter_c25_ramp", 24, 12, 1, synth.JITTER, 300, manual=(3, 25, 60, 2), dtm=255 * 60),
    # gradient scrolling through 0 with a large threshold: D = 128 nodes inside live stacks
    Case("off_gradient_c60", 48, 8, 1, synth.GRADIENT, 400, manual=(60, 60, 100, 1), dtm=255 * 100),
    Case("off_dim_flicker", 8, 4, 1, 0, 3 * 64, pattern=((5, 4), (0, 6), (9, 5), (1, 9), (0, 20), (3, 20)) * 3, dtm=255 * 40),
    # time_spanned an integer other than ref
    Case("off_time_2ref", 16, 8, 1, synth.JITTER, 120, crf=4, time_spanned=510.0),
    # DeltaT output and the other view modes run the general instantiation of px_offset
    Case("off_jitter_deltat", 24, 12, 1, synth.JITTER, 150, manual=(12, 12, 8, 1), dtm=255 * 8, time_mode=O.TIME_DELTA_T),
    Case("off_view_sae", 16, 8, 1, synth.JITTER, 80, crf=3, view_mode=O.VIEW_SAE),
]


def _run(case, check_state_every=0):
    sim_py.lib().sim_set_entry(4)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    sa = _SimAdapter(case)
    cases.configure(sa, case)
    total = 0
    step = 100
    for f0 in range(0, case.n_frames, step):
        n = min(step, case.n_frames - f0)
        frames = case.frames(f0, n)
        for k in range(n):
            f = f0 + k
            ev_o, _ = ov.integrate_matrix(frames[k], case.time)
            ev_s = sa.s.integrate(frames[k], case.time)
            assert len(ev_o) == len(ev_s), f"frame {f}: {len(ev_o)} vs {len(ev_s)} events"
            assert ev_o.tobytes() == ev_s.tobytes(), f"frame {f}: event streams differ"
            assert np.array_equal(ov.running_intensities(), sa.s.running()), f"frame {f}: display bytes differ"
            total += len(ev_o)
            if check_state_every and (f % check_state_every == 0 or f == case.n_frames - 1):
                for i in range(case.w * case.h * case.c):
                    a = cases.canonical_oracle_px(ov.px(i))
                    b = cases.canonical(sa.s.px(i))
                    assert a == b, f"frame {f} pixel {i}: state differs\noracle {a}\nsim    {b}"
    assert sa.s.err == 0
    assert total > 0
    return sa


@pytest.mark.parametrize("case", LONG, ids=lambda c: c.name)
def test_offset_form_long_runs(case):
    sa = _run(case, check_state_every=37)
    assert sa.s.form == 1, "the case was meant to run in offset form"


def test_eligibility_of_the_shared_cases():
    """Which of the shared cases the offset form serves (the rest fall back to the eager form): Collapse, integral time."""
    sim_py.lib().sim_set_entry(4)
    served = {}
    for case in cases.CASES:
        if case.initial_d or case.roi is not None:
            continue
        sa = _SimAdapter(case)
        cases.configure(sa, case)
        sa.s.integrate(case.frames(0, 1)[0], case.time)
        served[case.name] = sa.s.form
    assert served["cfg5_static_collapse"] == 1 and served["cfg2_rgb_noise_crf3"] == 1 and served["cfg3_jitter_c10"] == 1
    assert served["cfg5_static_normal"] == 0 and served["time_spanned_fraction"] == 0 and served["dark_params_jitter"] == 0


def test_a_frame_touches_one_level_record_whatever_the_depth():
    """The point of the form: on the long-integration workload the level records read + written per pixel-frame stay
    near one, while the stacks are 2..11 nodes deep (the eager form reads and writes every live level)."""
    case = Case("traffic", 16, 8, 1, synth.STATIC_BLIPS, 1200, crf=3, ref=256, dtm=1 << 20)
    sa = _run(case)
    loads, stores, pxf = sa.s.rec_traffic()
    depth = np.mean([sa.s.px(i)["length"] for i in range(case.w * case.h)])
    assert depth > 3.5
    assert loads / pxf < 0.75 and stores / pxf < 1.05, (loads / pxf, stores / pxf)


def _random_frames(rng, n, h, w, c):
    """Per pixel: stretches of a held value (1..120 frames), optionally with a small jitter, now and then 0 or 255 —
    changes, long integrations, zero-integration nodes and Δt_max pops in every mixture."""
    out = np.empty((n, h, w, c), dtype=np.uint8)
    for idx in np.ndindex(h, w, c):
        f = 0
        seq = np.empty(n, dtype=np.int32)
        while f < n:
            hold = int(rng.integers(1, 120))
            base = int(rng.choice([0, 1, 2, 5, 37, 128, 200, 254, 255, int(rng.integers(0, 256))]))
            amp = int(rng.choice([0, 0, 1, 3, 9]))
            seg = base + (rng.integers(-amp, amp + 1, hold) if amp else np.zeros(hold, dtype=np.int64))
            seq[f:f + hold] = seg[:n - f]
            f += hold
        out[(slice(None),) + idx] = np.clip(seq, 0, 255)
    return out


@pytest.mark.parametrize("seed", range(24))
def test_offset_form_random_parameters_and_scenes(seed):
    rng = np.random.default_rng(1000 + seed)
    ref = int(rng.choice([255, 256, 1000, 17]))
    mult = int(rng.choice([1, 2, 3, 5, 8, 13, 30, 64, 200]))
    c_base = int(rng.choice([0, 1, 4, 10, 25]))
    c_max = c_base + int(rng.choice([0, 0, 5, 30]))
    vel = int(rng.choice([1, 2, 5]))
    time_mode = O.TIME_DELTA_T if seed % 5 == 4 else None
    case = Case(f"rand{seed}", 6, 4, 1 if seed % 3 else 3, 0, 500, manual=(c_base, c_max, mult, vel), ref=ref, dtm=ref * mult, time_mode=time_mode)
    frames = _random_frames(rng, case.n_frames, case.h, case.w, case.c)
    sim_py.lib().sim_set_entry(4)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    sa = _SimAdapter(case)
    cases.configure(sa, case)
    for f in range(case.n_frames):
        ev_o, _ = ov.integrate_matrix(frames[f], case.time)
        ev_s = sa.s.integrate(frames[f], case.time)
        assert ev_o.tobytes() == ev_s.tobytes(), f"frame {f}: event streams differ"
        assert np.array_equal(ov.running_intensities(), sa.s.running()), f"frame {f}: display bytes differ"
        if f % 61 == 0 or f == case.n_frames - 1:
            for i in range(case.w * case.h * case.c):
                assert cases.canonical_oracle_px(ov.px(i)) == cases.canonical(sa.s.px(i)), f"frame {f} pixel {i}"
    assert sa.s.err == 0 and sa.s.form == 1
