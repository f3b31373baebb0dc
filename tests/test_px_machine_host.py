"""The product's per-pixel state machine (px_machine.cuh, compiled for the host by tests/host_sim)
against the oracle: events, display bytes and final state, on every shared case.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import cases
from tests.sim_py import SimVideo


class _SimAdapter:
    """Gives SimVideo the mirrored Video interface that cases.configure() drives."""

    def __init__(self, case):
        depth = 31
        self.s = SimVideo(case.w, case.h, case.c, depth)
        self.in_interval = 1

    def chunk_rows(self, n):
        pass

    def time_parameters(self, tps, ref, dtm, time_mode):
        if time_mode is not None:
            self.s.abs_time = int(time_mode == O.TIME_ABSOLUTE_T)
        if dtm < ref:
            return False
        self.s.ref, self.s.dtm = ref, dtm
        self.s.force_display()
        return True

    def write_out(self, time_mode, multi_mode):
        self.s.collapse = 1 if multi_mode is None else int(multi_mode == O.MULTI_COLLAPSE)
        if time_mode is not None:
            self.s.abs_time = int(time_mode == O.TIME_ABSOLUTE_T)

    def update_crf(self, crf):
        p = O.crf_parameters(crf, self.s.w, self.s.h)
        self.s.c_max, self.s.vel = p.c_thresh_max, p.c_increase_velocity
        self.s.reset_c(p.c_thresh_baseline)

    def update_quality_manual(self, c_base, c_max, dtm_mult, velocity, radius):
        self.s.c_max, self.s.vel = c_max, velocity
        self.s.dtm = dtm_mult * self.s.ref
        self.s.force_display()
        self.s.reset_c(c_base)

    def set_view_mode(self, m):
        self.s.view = m
        self.s.force_display()

    def set_in_interval_count(self, n):
        self.in_interval = n


@pytest.mark.parametrize("entry", [0, 1, 2, 3, 4], ids=["px_step", "px_frame", "px_step_deferred_deep_walk", "px_step_plain", "px_offset"])
@pytest.mark.parametrize("case", [c for c in cases.CASES if not c.initial_d and c.roi is None], ids=lambda c: c.name)
def test_machine_matches_oracle(case, entry):
    from tests import sim_py

    sim_py.lib().sim_set_entry(entry)  # the general state machine (with the record walk), and the short path in front of it
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    sa = _SimAdapter(case)
    cases.configure(sa, case)
    frames = case.frames()
    total = 0
    for f in range(case.n_frames):
        ev_o, _ = ov.integrate_matrix(frames[f], case.time)
        ev_s = sa.s.integrate(frames[f], case.time)
        assert len(ev_o) == len(ev_s), f"frame {f}: {len(ev_o)} vs {len(ev_s)} events"
        assert ev_o.tobytes() == ev_s.tobytes(), f"frame {f}: event streams differ"
        assert np.array_equal(ov.running_intensities(), sa.s.running()), f"frame {f}: display bytes differ"
        total += len(ev_o)
    assert sa.s.err == 0
    assert total > 0 or case.constant == 0
    for i in range(case.w * case.h * case.c):
        a = cases.canonical_oracle_px(ov.px(i))
        b = cases.canonical(sa.s.px(i))
        assert a == b, f"pixel {i}: state differs\noracle {a}\nsim    {b}"
