"""The exact-by-construction shortcuts of px_machine.cuh (host build): the multiply-by-reciprocal
`u / ref_time`, the f32-estimated Intensity display byte (with the estimate perturbed by a few ulp to
stand for the device's approximate divide) and the skip of unchanged display bytes.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import cases, synth
from tests.sim_py import SimVideo, lib as sim_lib
from tests.test_px_machine_host import _SimAdapter


def test_div_ref_is_exact():
    L = sim_lib()
    rng = np.random.default_rng(1)
    refs = [1, 2, 3, 5, 7, 255, 256, 257, 1000, 6120, 65535, 65536, 1 << 20, (1 << 31) - 1, 1 << 31, (1 << 32) - 1]
    refs += [int(x) for x in rng.integers(1, 1 << 32, 40)]
    for ref in refs:
        us = [0, 1, ref - 1, ref, ref + 1, 2 * ref - 1, 2 * ref, (1 << 32) - 1, (1 << 32) - ref, ((1 << 32) // ref) * ref - 1,
              ((1 << 32) // ref) * ref]
        us += [int(x) for x in rng.integers(0, 1 << 32, 300)]
        us += [int(k) * ref for k in rng.integers(0, max(1, (1 << 32) // ref), 50)]
        for u in us:
            u &= 0xFFFFFFFF
            assert L.sim_div_ref(u, ref) == u // ref, (u, ref)


@pytest.mark.parametrize("ulps", [0, -3, 3, -8, 8])
def test_intensity_display_byte_matches_f64_reference(ulps):
    """Against the oracle's f64 restatement of event_to_intensity * tpf (scale_intensity.rs:262-270, :58-68)."""
    L = sim_lib()
    OL = O.lib()
    L.sim_set_fast_div_ulps(ulps)
    try:
        rng = np.random.default_rng(70 + ulps)
        trials = []
        for ref in (1, 3, 100, 255, 256, 1000, 6120):
            # exact-integer and near-integer quotients: t divides 2^d * ref, or misses by one
            for d in range(0, 20):
                for k in (1, 2, 3, 5, 64, 100, 127, 128, 200, 255, 256, 257):
                    n = (1 << d) * ref
                    if n % k == 0:
                        t = n // k
                        for tt in (t - 1, t, t + 1):
                            if 0 <= tt < (1 << 32):
                                trials.append((d, tt, ref))
            for d in (0, 1, 6, 7, 8, 30, 31, 32, 40, 64, 126, 127, 128, 129, 200, 255):
                for t in (0, 1, 2, 254, 255, 256, 65535, (1 << 24) + 1, (1 << 32) - 1):
                    trials.append((d, t, ref))
            ds = rng.integers(0, 24, 3000)
            ts = rng.integers(1, 1 << 22, 3000)
            trials += [(int(d), int(t), ref) for d, t in zip(ds, ts)]
            # the values the transcoder really produces: t = trunc(255 * 2^d / I)-like
            for inten in range(1, 256):
                d = inten.bit_length() - 1
                for m in (1, 2, 3):
                    trials.append((d + m - 1, int(ref * m * (1 << d) / inten), ref))
        # every exact quotient k = 1..256 for the default ref, and huge t / ref combinations
        for d in range(0, 14):
            for k in range(1, 257):
                n = (1 << d) * 255
                if n % k == 0:
                    trials += [(d, n // k + dt, 255) for dt in (-1, 0, 1) if n // k + dt >= 0]
        for ref in (65536, (1 << 31) + 1, (1 << 32) - 1, 3000000007 % (1 << 32)):
            for d in range(0, 40):
                for t in ((ref << d) // 255 % (1 << 32), (1 << 32) - 1, (1 << 31), ref, ref - 1, max(1, (ref << d) >> 8) % (1 << 32)):
                    trials.append((d, int(t), ref))
            trials += [(int(d), int(t), ref) for d, t in zip(rng.integers(0, 40, 500), rng.integers(1, 1 << 32, 500))]
        for d, t, ref in trials:
            want = OL.oracle_get_frame_value_u8(d, t, float(ref), 0.0, 7650, 0, 0, 0)
            got = L.sim_frame_value_intensity(d, t, ref)
            assert got == want, (d, t, ref, got, want)
    finally:
        L.sim_set_fast_div_ulps(0)


def test_display_skip_survives_parameter_changes_mid_run():
    """View mode and time parameters change between frames: bytes whose root best event did not
    change must still be recomputed once (PxParams::display == 2)."""
    case = cases.Case("static_modes", 24, 10, 1, synth.STATIC_BLIPS, 60, crf=3, ref=255, dtm=255 * 64)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    sa = _SimAdapter(case)
    cases.configure(sa, case)
    frames = case.frames()
    for f in range(case.n_frames):
        if f == 20:
            ov.set_view_mode(O.VIEW_DELTA_T)
            sa.set_view_mode(O.VIEW_DELTA_T)
        if f == 30:
            ov.set_view_mode(O.VIEW_INTENSITY)
            sa.set_view_mode(O.VIEW_INTENSITY)
        if f == 40:
            assert ov.time_parameters(255 * 30, 255, 255 * 32, None)
            assert sa.time_parameters(255 * 30, 255, 255 * 32, None)
            ov.set_view_mode(O.VIEW_D)
            sa.set_view_mode(O.VIEW_D)
        if f == 50:
            ov.set_view_mode(O.VIEW_SAE)
            sa.set_view_mode(O.VIEW_SAE)
        ev_o, _ = ov.integrate_matrix(frames[f], case.time)
        ev_s = sa.s.integrate(frames[f], case.time)
        assert ev_o.tobytes() == ev_s.tobytes(), f"frame {f}"
        assert np.array_equal(ov.running_intensities(), sa.s.running()), f"frame {f}: display bytes differ"


@pytest.mark.parametrize("seed", range(6))
def test_random_parameter_fuzz(seed):
    """Random geometry-free fuzz of the state machine: ref, dtm multiple, c range, velocity, modes."""
    rng = np.random.default_rng(100 + seed)
    ref = int(rng.choice([1, 7, 100, 255, 256, 1000]))
    mult = int(rng.choice([1, 2, 3, 5, 16, 100]))
    c0 = int(rng.integers(0, 12))
    c1 = c0 + int(rng.integers(0, 8))
    vel = int(rng.integers(1, 9))
    kind = int(rng.choice([synth.NOISE, synth.JITTER, synth.STATIC_BLIPS, synth.GRADIENT]))
    case = cases.Case(f"fuzz{seed}", 32, 8, 1, kind, 90, seed=seed, manual=(c0, c1, mult, vel), ref=ref, dtm=ref * mult,
                      multi_mode=int(rng.integers(0, 2)), time_mode=int(rng.integers(0, 2)))
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    sa = _SimAdapter(case)
    cases.configure(sa, case)
    frames = case.frames()
    for f in range(case.n_frames):
        ev_o, _ = ov.integrate_matrix(frames[f], case.time)
        ev_s = sa.s.integrate(frames[f], case.time)
        assert ev_o.tobytes() == ev_s.tobytes(), f"frame {f}"
        assert np.array_equal(ov.running_intensities(), sa.s.running()), f"frame {f}"
    assert sa.s.err == 0


def test_gray_conversion_shortcut_is_exact_for_every_colour():
    """gray_math.h: 2^24 fixed point, the diagonal table and the f64 fallback against the reference's f64 expression
    for all 16.7 M (c0, c1, c2); the same host build against the oracle's handle_color on a sample including the
    colours whose weighted sum is an integer up to rounding."""
    import ctypes as C

    L = sim_lib()
    n_fast = C.c_uint64()
    assert L.sim_gray_check(C.byref(n_fast)) == 0
    assert n_fast.value > 0.998 * (1 << 24)  # the shortcut decides nearly everything (16 774 of 16.7 M colours fall back)
    rng = np.random.default_rng(5)
    cols = np.concatenate([rng.integers(0, 256, (4000, 3)), np.repeat(np.arange(256)[:, None], 3, axis=1),
                           np.array([[255, 255, 255], [0, 0, 0], [250, 0, 0], [0, 255, 0], [1, 1, 0]])]).astype(np.uint8)
    want = O.handle_color(cols.reshape(1, -1, 3)).reshape(-1)
    got = np.array([L.sim_gray_of(int(a), int(b), int(c)) for a, b, c in cols], dtype=np.uint8)
    assert np.array_equal(got, want)


def test_raw_pack4_matches_the_oracle_encoder():
    """raw_pack.h (what a thread of raw_encode_kernel runs: four records -> 36 / 44 wire bytes on registers) against the oracle's
    serialiser (RawOutput::ingest_event, raw/stream.rs:100-120), both record sizes, extreme field values included."""
    L = sim_lib()
    rng = np.random.default_rng(11)
    n = 4096
    ev = np.zeros(n, dtype=O.EVENT_DTYPE)
    ev["x"] = rng.integers(0, 65536, n)
    ev["y"] = rng.integers(0, 65536, n)
    ev["c"] = rng.integers(0, 3, n)
    ev["d"] = rng.integers(0, 256, n)
    ev["t"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    ev[:4] = [(0x0102, 0x0304, 2, 7, 0, 0x0A0B0C0D), (65535, 0, 0, 255, 0, 0xFFFFFFFF), (0, 65535, 1, 0, 0, 0), (1, 2, 2, 0xFE, 0, 0x80000001)]
    words = np.frombuffer(ev.tobytes(), dtype=np.uint32).copy()
    for channels, esize in ((3, 11), (1, 9)):
        e = ev.copy()
        if channels == 1:
            e["c"] = O.C_NONE
            words = np.frombuffer(e.tobytes(), dtype=np.uint32).copy()
        out = np.zeros(n * esize, dtype=np.uint8)
        L.sim_raw_pack(words.ctypes.data, n, esize, out.ctypes.data)
        assert out.tobytes() == O.raw_encode(e, channels), f"esize {esize}"
