"""Shared parity cases: one description drives the oracle, the host simulation and the GPU path.

Each case is a small instance of a BASELINE.json config family or an edge case
the reference's own tests exercise (Δt_max pops, D=128 zero events, Collapse/Normal, DeltaT/AbsoluteT).
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from oracle import oracle_py as O
from tests import synth


@dataclass
class Case:
    name: str
    w: int
    h: int
    c: int
    kind: int
    n_frames: int
    seed: int = 0xADDE5
    crf: Optional[int] = None            # update_crf(n); None = API default (c_thresh 10 fixed)
    manual: Optional[tuple] = None       # (c_base, c_max, dtm_mult, velocity)
    ref: int = 255
    dtm: int = 7650
    time_mode: Optional[int] = None      # None = keep default AbsoluteT
    multi_mode: Optional[int] = None     # None = default Collapse
    view_mode: int = O.VIEW_INTENSITY
    chunk_rows: int = 1
    time_spanned: Optional[float] = None # default = ref
    initial_d: bool = False              # zero in_interval_count first (adder-viz restart path)
    roi: Optional[tuple] = None          # (frame_idx, x0, y0, x1, y1, value)
    constant: Optional[int] = None       # constant-intensity frames instead of `kind`
    pattern: Optional[tuple] = None      # ((value, n_frames), ...) whole-plane constant segments instead of `kind`

    def frames(self, f0=0, n=None):
        n = self.n_frames if n is None else n
        if self.constant is not None:
            return np.full((n, self.h, self.w, self.c), self.constant, dtype=np.uint8)
        if self.pattern is not None:
            vals = np.concatenate([np.full(k, val, dtype=np.uint8) for val, k in self.pattern])
            assert len(vals) >= f0 + n
            return np.broadcast_to(vals[f0:f0 + n, None, None, None], (n, self.h, self.w, self.c)).copy()
        return synth.frames(self.kind, self.seed, f0, n, self.w, self.h, self.c)

    @property
    def time(self):
        return float(self.ref if self.time_spanned is None else self.time_spanned)


def configure(v, case: Case):
    """Apply a case to any object with the mirrored Video interface (oracle_py.Video or the GPU Video)."""
    if case.chunk_rows != 1:
        v.chunk_rows(case.chunk_rows)
    assert v.time_parameters(case.ref * 30, case.ref, case.dtm, case.time_mode)
    if case.multi_mode is not None or case.time_mode is not None:
        v.write_out(case.time_mode, case.multi_mode)
    if case.crf is not None:
        v.update_crf(case.crf)
    if case.manual is not None:
        dtm_keep = case.dtm
        v.update_quality_manual(case.manual[0], case.manual[1], case.manual[2], case.manual[3], 0.0)
        assert case.manual[2] * case.ref == dtm_keep, "manual dtm multiplier must agree with case.dtm"
    v.set_view_mode(case.view_mode)
    if case.initial_d:
        v.set_in_interval_count(0)


CASES = [
    # BASELINE config 1 family: gradient, API defaults (c_thresh 10), dtm = ref
    Case("cfg1_gradient_dtm_eq_ref", 64, 48, 1, synth.GRADIENT, 30, dtm=255),
    # config 2 family: RGB noise, crf 3
    Case("cfg2_rgb_noise_crf3", 40, 24, 3, synth.NOISE, 40, crf=3),
    Case("cfg2_rgb_noise_api_default", 40, 24, 3, synth.NOISE, 24),
    # config 3 family: jitter, c sweep
    Case("cfg3_jitter_c0", 48, 32, 1, synth.JITTER, 70, manual=(0, 0, 30, 1)),
    Case("cfg3_jitter_c5", 48, 32, 1, synth.JITTER, 70, manual=(5, 5, 30, 1)),
    Case("cfg3_jitter_c10", 48, 32, 1, synth.JITTER, 100, manual=(10, 10, 30, 1)),
    # config 5 family: long integration, static + blips, ref 256 dtm 2^20, both multi modes
    Case("cfg5_static_collapse", 32, 16, 1, synth.STATIC_BLIPS, 160, crf=3, ref=256, dtm=1 << 20),
    Case("cfg5_static_normal", 32, 16, 1, synth.STATIC_BLIPS, 160, crf=3, ref=256, dtm=1 << 20, multi_mode=O.MULTI_NORMAL),
    # the reference's `dark` test parameters: crf 0, DeltaT, Normal, dtm 6120
    Case("dark_params_jitter", 50, 20, 1, synth.JITTER, 60, crf=0, dtm=6120, time_mode=O.TIME_DELTA_T, multi_mode=O.MULTI_NORMAL),
    # Δt_max pops on constant input (dtm = 4 ref so they come often), both multi modes, both time modes
    Case("const100_dtm4_collapse", 16, 8, 1, 0, 40, constant=100, dtm=255 * 4),
    Case("const100_dtm4_normal_deltat", 16, 8, 1, 0, 40, constant=100, dtm=255 * 4, multi_mode=O.MULTI_NORMAL, time_mode=O.TIME_DELTA_T),
    Case("const0_no_events", 16, 8, 1, 0, 12, constant=0, dtm=255 * 4),
    Case("zero_then_bright", 16, 8, 1, 0, 60, pattern=((0, 7), (200, 9), (0, 3), (5, 11), (0, 10), (8, 10), (255, 10)), dtm=255 * 4),
    Case("zero_then_bright_normal_deltat", 16, 8, 1, 0, 60, pattern=((0, 7), (200, 9), (0, 3), (5, 11), (0, 10), (8, 10), (255, 10)),
         dtm=255 * 4, multi_mode=O.MULTI_NORMAL, time_mode=O.TIME_DELTA_T),
    Case("dim_flicker_c10", 16, 8, 1, 0, 64, pattern=((5, 4), (0, 6), (9, 5), (1, 9), (0, 20), (3, 20)), dtm=255 * 6),
    Case("const1_dtm8", 16, 8, 1, 0, 40, constant=1, dtm=255 * 8, multi_mode=O.MULTI_NORMAL),
    Case("const255_dtm8", 16, 8, 3, 0, 40, constant=255, dtm=255 * 8),
    # jitter with frequent dtm pops in both multi modes (pop_best after popped_dtm -> D_EMPTY path)
    Case("jitter_dtm4_collapse", 48, 16, 1, synth.JITTER, 80, manual=(12, 12, 4, 1), dtm=255 * 4),
    Case("jitter_dtm4_normal", 48, 16, 1, synth.JITTER, 80, manual=(12, 12, 4, 1), dtm=255 * 4, multi_mode=O.MULTI_NORMAL),
    Case("jitter_dtm4_collapse_deltat", 48, 16, 1, synth.JITTER, 80, manual=(12, 12, 4, 1), dtm=255 * 4, time_mode=O.TIME_DELTA_T),
    Case("blips_dtm16_collapse", 48, 16, 3, synth.STATIC_BLIPS, 120, crf=5, dtm=255 * 16),
    # gradient scrolls through 0 with large c_thresh: zero-integration nodes inside live stacks
    Case("gradient_c40", 64, 16, 1, synth.GRADIENT, 90, manual=(40, 40, 30, 1)),
    Case("gradient_c40_normal", 64, 16, 1, synth.GRADIENT, 90, manual=(40, 40, 30, 1), multi_mode=O.MULTI_NORMAL),
    # ragged geometry: plane not a multiple of the tile, chunk_rows not dividing H, odd channel count
    Case("ragged_37x13x3_chunk4", 37, 13, 3, synth.NOISE, 12, crf=3, chunk_rows=4),
    Case("ragged_1x1", 1, 1, 1, synth.NOISE, 50, crf=3),
    Case("ragged_3x700_chunk64", 3, 700, 1, synth.JITTER, 10, crf=3, chunk_rows=64),
    Case("ragged_300x2x2_chunk5", 300, 2, 2, synth.NOISE, 10, crf=2, chunk_rows=5),
    # view modes of the display byte
    Case("view_d", 32, 16, 1, synth.JITTER, 40, crf=3, view_mode=O.VIEW_D),
    Case("view_delta_t", 32, 16, 1, synth.JITTER, 40, crf=3, view_mode=O.VIEW_DELTA_T),
    Case("view_sae", 32, 16, 1, synth.JITTER, 40, crf=3, view_mode=O.VIEW_SAE),
    # time_spanned different from ref (Prophesee's first frames pass other spans), c ramp velocity > 1
    Case("time_spanned_2ref", 32, 16, 1, synth.JITTER, 40, crf=4, time_spanned=510.0),
    Case("time_spanned_fraction", 32, 16, 1, synth.JITTER, 40, crf=4, time_spanned=100.5),
    # restart path: in_interval_count = 0 -> set_initial_d
    Case("initial_d", 32, 16, 1, synth.NOISE, 20, crf=3, initial_d=True),
    # ROI write between frames
    Case("roi_rect", 32, 16, 3, synth.JITTER, 40, crf=6, roi=(10, 4, 3, 20, 9, 2)),
]
CASES_BY_NAME = {c.name: c for c in CASES}


def canonical_oracle_px(px) -> dict:
    """Oracle pixel -> comparable dict.  The d of an all-zero tail node is not compared: the
    reference rewrites it at the start of every integrate (event_pixel_tree.rs:332-335) before it can
    be used, and the kernel applies that rewrite lazily."""
    st = dict(last_fired_t=px.last_fired_t, base_val=px.base_val, c_thresh=px.c_thresh,
              c_increase_counter=px.c_increase_counter, length=px.length, popped_dtm=px.popped_dtm, nodes=[])
    nodes = px.heap if px.heap else px.inl
    for k in range(px.length):
        n = nodes[k]
        st["nodes"].append(dict(integration=n.integration, delta_t=n.delta_t, d=n.d, has_best=n.has_best,
                                best_d=n.best_d if n.has_best else 0, best_delta_t=n.best_delta_t if n.has_best else 0.0))
    return canonical(st)


def canonical(st: dict) -> dict:
    out = {k: st[k] for k in ("last_fired_t", "base_val", "c_thresh", "c_increase_counter", "length", "popped_dtm")}
    out["nodes"] = []
    for k, n in enumerate(st["nodes"]):
        m = dict(integration=n["integration"], delta_t=n["delta_t"], d=n["d"], has_best=int(n["has_best"]),
                 best_d=n["best_d"] if n["has_best"] else 0, best_delta_t=n["best_delta_t"] if n["has_best"] else 0.0)
        if k == st["length"] - 1 and m["integration"] == 0.0 and m["delta_t"] == 0.0:
            m["d"] = None
        out["nodes"].append(m)
    return out
