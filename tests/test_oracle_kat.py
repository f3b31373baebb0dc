"""Known-answer tests of the oracle's pixel state machine.

Each test restates one of the reference's 13 unit tests,
adder-codec-rs/src/transcoder/event_pixel_tree.rs:534-1259, with the same scripted
integrate/pop sequence and the same asserted values.  This is what pins the oracle.
"""
import math
import struct

import numpy as np
import pytest

from oracle.oracle_py import (
    MODE_CONTINUOUS as Continuous,
    MODE_FRAME_PERFECT as FramePerfect,
    MULTI_COLLAPSE,
    MULTI_NORMAL,
    TIME_ABSOLUTE_T,
    TIME_DELTA_T,
    PixelArena,
)

EPS = float(np.finfo(np.float32).eps)


def f32_slack(a, b):  # event_pixel_tree.rs:1005-1011
    b = np.float32(b)
    return float(b - np.float32(EPS)) <= a <= float(b + np.float32(EPS))


def ulps(a, b):
    ia = struct.unpack("<i", struct.pack("<f", a))[0]
    ib = struct.unpack("<i", struct.pack("<f", float(np.float32(b))))[0]
    return abs(ia - ib)


def integ(tree, i, t, mode, dtm, ref, multi=MULTI_NORMAL):
    tree.integrate(i, t, mode, dtm, ref, 0, 255, multi)


def make_tree():  # :541-639
    dtm = 10_000
    tree = PixelArena(100.0)
    tree.time_mode(TIME_DELTA_T)
    assert tree.node(0).d == 6
    integ(tree, 100.0, 20.0, Continuous, dtm, 20)
    n0 = tree.node(0)
    assert n0.has_best and n0.best_d == 6 and int(n0.best_delta_t) == 12
    assert n0.d == 7 and f32_slack(n0.integration, 100.0) and f32_slack(n0.delta_t, 20.0) and n0.alt
    n1 = tree.node(1)
    assert not n1.has_best and n1.d == 6 and n1.integration == 36.0
    assert ulps(n1.delta_t, 7.2) <= 2

    integ(tree, 100.0, 20.0, Continuous, dtm, 20)
    n0, n1, n2 = tree.node(0), tree.node(1), tree.node(2)
    assert n0.best_d == 7 and ulps(n0.best_delta_t, 25.6) <= 1
    assert n0.d == 8 and f32_slack(n0.integration, 200.0) and f32_slack(n0.delta_t, 40.0) and n0.alt
    assert n1.d == 7 and f32_slack(n1.integration, 72.0) and ulps(n1.delta_t, 14.4) <= 1
    assert n1.best_d == 6 and ulps(n1.best_delta_t, 12.8) <= 2 and n1.alt
    assert n2.d == 6 and not n2.has_best and not n2.alt and f32_slack(n2.integration, 8.0)
    assert abs(n2.delta_t - 1.6) <= 0.2e-5
    return tree


def make_tree2():  # :641-709
    dtm = 10_000
    tree = make_tree()
    integ(tree, 30.0, 34.0, Continuous, dtm, 34)
    n0, n1, n2 = tree.node(0), tree.node(1), tree.node(2)
    assert n0.d == 8 and f32_slack(n0.integration, 230.0) and f32_slack(n0.delta_t, 74.0)
    assert n1.d == 7 and f32_slack(n1.integration, 102.0) and f32_slack(n1.delta_t, 48.4)
    assert n2.d == 6 and f32_slack(n2.integration, 38.0) and f32_slack(n2.delta_t, 35.6)
    integ(tree, 26.0, 34.0, Continuous, dtm, 34)
    n0, n1 = tree.node(0), tree.node(1)
    assert n0.d == 9 and f32_slack(n0.integration, 256.0) and f32_slack(n0.delta_t, 108.0)
    assert n0.best_d == 8 and n0.best_delta_t == 108.0
    assert n1.d == 4 and f32_slack(n1.integration, 0.0) and f32_slack(n1.delta_t, 0.0)
    assert not n1.has_best and not n1.alt
    return tree


def test_make_tree():
    make_tree()


def test_make_tree2():
    make_tree2()


def test_pop_best_states():  # :721-741
    tree = make_tree()
    events = tree.pop_best_events(Continuous, MULTI_NORMAL, 20, 0.0)
    assert events == [(7, 25), (6, 12)]
    n0 = tree.node(0)
    assert n0.d == 6 and f32_slack(n0.integration, 8.0) and abs(n0.delta_t - 1.6) <= 0.2e-5


def test_pop_best_states2():  # :743-755
    tree = make_tree2()
    events = tree.pop_best_events(Continuous, MULTI_NORMAL, 34, 0.0)
    assert events == [(8, 108)]
    n0 = tree.node(0)
    assert n0.d == 4 and f32_slack(n0.integration, 0.0) and f32_slack(n0.delta_t, 0.0)


def test_d_max():  # :757-794
    dtm = 100_000_000
    big = float(np.float32(2.0**126))
    tree = PixelArena(big)
    integ(tree, float(np.float32(big) + np.float32(5.0)), 100_000.0, Continuous, dtm, 100_000)
    assert tree.need_to_pop_top
    events = tree.pop_best_events(Continuous, MULTI_NORMAL, 100_000, 0.0)
    assert not tree.need_to_pop_top
    assert events == [(126, 100_000)]
    assert f32_slack(tree.node(0).integration, 0.0)


def test_dtm():  # :796-834
    dtm = 240_000
    tree = PixelArena(245.0)
    for _ in range(48):
        integ(tree, 245.0, 5_000.0, FramePerfect, dtm, 5_000)
    assert tree.need_to_pop_top
    tree.pop_top_event(245.0, FramePerfect, 5_000)
    assert not tree.need_to_pop_top
    assert tree.node(0).delta_t == 70_000.0


def test_new_dtm():  # :836-925
    dtm = 2_000
    tree = PixelArena(245.0)
    integ(tree, 245.0, 1_000.0, FramePerfect, dtm, 5_000)
    assert not tree.need_to_pop_top
    integ(tree, 245.0, 1_000.0, FramePerfect, dtm, 5_000)
    assert tree.need_to_pop_top
    tree.pop_top_event(245.0, FramePerfect, 5_000)
    assert not tree.need_to_pop_top
    for _ in range(48):
        integ(tree, 245.0, 1_000.0, FramePerfect, dtm, 5_000)
    assert not tree.need_to_pop_top
    assert tree.node(0).delta_t == 48000.0
    tree.pop_best_events(FramePerfect, MULTI_COLLAPSE, 5_000, 0.0)
    integ(tree, 600.0, 3_000.0, FramePerfect, dtm, 5_000)
    assert tree.need_to_pop_top


def test_big_integration():  # :927-966
    dtm = 1_000_000
    tree = PixelArena(146.0)
    integ(tree, 146.0, 2_000.0, Continuous, dtm, 2_000)
    integ(tree, float(np.float32(2_790.863)), 38231.0, Continuous, dtm, 38231)
    head = tree.node(0)
    assert head.integration == float(np.float32(2_790.863) + np.float32(146.0))
    assert head.delta_t == 38231.0 + 2_000.0
    assert head.best_d == head.d - 1


def test_big_integration2():  # :968-1003
    dtm = 10_000_000
    tree = PixelArena(255.0)
    for _ in range(100_000):
        integ(tree, 255.0, 2_000.0, Continuous, dtm, 2_000)
        if tree.need_to_pop_top:
            break
    head = tree.node(0)
    assert head.integration == 1.275e6
    assert head.delta_t == float(dtm)
    assert head.best_d == head.d - 1


def test_paper_example():  # :1021-1060
    dtm = 10_000
    tree = PixelArena(101.0)
    assert tree.node(0).d == 6
    integ(tree, 101.0, 20.0, Continuous, dtm, 20)
    assert tree.node(0).has_best
    integ(tree, 40.0, 30.0, Continuous, dtm, 30)
    assert tree.node(0).best_d == 7
    assert f32_slack(tree.node(1).delta_t, 9.75)


def _absolute_script(last):
    dtm = 10_000
    tree = PixelArena(101.0)
    assert tree.node(0).d == 6
    return tree, dtm


def test_absolute_mode_1():  # :1062-1126
    tree, dtm = _absolute_script(None)
    tree.time_mode(TIME_ABSOLUTE_T)
    integ(tree, 101.0, 20.0, Continuous, dtm, 20)
    assert tree.node(0).has_best
    integ(tree, 40.0, 30.0, Continuous, dtm, 30)
    integ(tree, 140.0, 30.0, Continuous, dtm, 30)
    integ(tree, 103.0, 30.0, Continuous, dtm, 30)
    events = tree.pop_best_events(Continuous, MULTI_COLLAPSE, 30, 0.0)
    assert events[0] == (8, 74)
    assert events[1] == (7, 110)


@pytest.mark.parametrize("time_mode,expect_t", [(TIME_DELTA_T, 1), (TIME_ABSOLUTE_T, 110)])
def test_set_d_continuous(time_mode, expect_t):  # :1128-1258 (delta and absolute variants)
    tree, dtm = _absolute_script(None)
    tree.time_mode(time_mode)
    integ(tree, 101.0, 20.0, Continuous, dtm, 20)
    assert tree.node(0).has_best
    integ(tree, 40.0, 30.0, Continuous, dtm, 30)
    integ(tree, 140.0, 30.0, Continuous, dtm, 30)
    integ(tree, 107.0, 30.0, Continuous, dtm, 30)
    tree.pop_best_events(Continuous, MULTI_COLLAPSE, 30, 0.0)
    ev = tree.set_d_for_continuous(10.0, 30)
    assert ev == (255, expect_t)


def test_get_d_from_intensity(oracle):
    """event_pixel_tree.rs:482-499: <1 -> 128, else floor(log2), clamped to 127."""
    L = oracle.lib()
    assert L.oracle_get_d_from_intensity(0.0) == 128
    assert L.oracle_get_d_from_intensity(0.999) == 128
    assert L.oracle_get_d_from_intensity(1.0) == 0
    for v in range(1, 256):
        assert L.oracle_get_d_from_intensity(float(v)) == int(math.floor(math.log2(v)))
    assert L.oracle_get_d_from_intensity(float(np.float32(2.0**127))) == 127
    assert L.oracle_get_d_from_intensity(float(np.float32(3.0e38))) == 127
    assert L.oracle_get_d_from_intensity(1.275e6) == 20


def test_crf_table(oracle):
    """rate_controller.rs:5-18, :55-70."""
    p = oracle.crf_parameters(3, 1920, 1080)
    assert (p.c_thresh_baseline, p.c_thresh_max, p.c_increase_velocity, p.feature_c_radius) == (2, 7, 7, 72)
    p = oracle.crf_parameters(0, 200, 50)
    assert (p.c_thresh_baseline, p.c_thresh_max, p.c_increase_velocity, p.feature_c_radius) == (0, 0, 10, 0)
    p = oracle.crf_parameters(9, 640, 480)
    assert (p.c_thresh_baseline, p.c_thresh_max, p.c_increase_velocity, p.feature_c_radius) == (15, 25, 1, 16)


def test_oracle_synth_frames_equal_the_numpy_generator():
    """oracle_synth_frame (the CPU bench arm's frame source) == tests/synth.py for every kind, including a row band."""
    from oracle import oracle_py as O
    from tests import synth

    for kind in (synth.GRADIENT, synth.NOISE, synth.JITTER, synth.STATIC_BLIPS):
        for (w, h, c) in ((37, 13, 3), (64, 9, 1)):
            want = synth.frame(kind, 0xADDE5, 7, w, h, c)
            assert np.array_equal(O.synth_frame(kind, 0xADDE5, 7, w, h, c), want)
            assert np.array_equal(O.synth_frame(kind, 0xADDE5, 7, w, 4, c, row0=5), want[5:9])
    assert O.host_threads() >= 1
