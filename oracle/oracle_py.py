"""ctypes binding of the CPU oracle (oracle/libadder_oracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/adder_oracle.h): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs — never by the
product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libadder_oracle.so")

# adder_event_t, include/adder_b200.h (12 bytes, little-endian)
EVENT_DTYPE = np.dtype(
    [("x", "<u2"), ("y", "<u2"), ("c", "u1"), ("d", "u1"), ("reserved", "<u2"), ("t", "<u4")]
)
assert EVENT_DTYPE.itemsize == 12

MODE_FRAME_PERFECT, MODE_CONTINUOUS = 0, 1
MULTI_NORMAL, MULTI_COLLAPSE = 0, 1
TIME_DELTA_T, TIME_ABSOLUTE_T, TIME_MIXED = 0, 1, 2
VIEW_INTENSITY, VIEW_D, VIEW_DELTA_T, VIEW_SAE = 0, 1, 2, 3
D_MAX, D_ZERO_INTEGRATION, D_EMPTY, C_NONE = 127, 128, 255, 0xFF


class Node(C.Structure):
    _fields_ = [
        ("integration", C.c_float),
        ("delta_t", C.c_float),
        ("best_delta_t", C.c_float),
        ("d", C.c_uint8),
        ("has_best", C.c_uint8),
        ("best_d", C.c_uint8),
        ("alt", C.c_uint8),
    ]


class Event(C.Structure):
    _fields_ = [
        ("x", C.c_uint16),
        ("y", C.c_uint16),
        ("c", C.c_uint8),
        ("d", C.c_uint8),
        ("reserved", C.c_uint16),
        ("t", C.c_uint32),
    ]


class Px(C.Structure):
    _fields_ = [
        ("x", C.c_uint16),
        ("y", C.c_uint16),
        ("c", C.c_uint8),
        ("time_mode", C.c_uint8),
        ("base_val", C.c_uint8),
        ("need_to_pop_top", C.c_uint8),
        ("c_thresh", C.c_uint8),
        ("c_increase_counter", C.c_uint8),
        ("dtm_reached", C.c_uint8),
        ("popped_dtm", C.c_uint8),
        ("last_fired_t", C.c_float),
        ("running_t", C.c_float),
        ("length", C.c_uint32),
        ("arena_len", C.c_uint32),
        ("arena_cap", C.c_uint32),
        ("heap", C.POINTER(Node)),
        ("inl", Node * 6),
    ]


class CrfParameters(C.Structure):
    _fields_ = [
        ("c_thresh_baseline", C.c_uint8),
        ("c_thresh_max", C.c_uint8),
        ("c_increase_velocity", C.c_uint8),
        ("reserved", C.c_uint8),
        ("feature_c_radius", C.c_uint16),
        ("reserved2", C.c_uint16),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("adder_oracle.c", "framer_oracle.c", "adder_oracle.h", "Makefile")]
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, u8, u16, u32, f32, i32 = C.c_void_p, C.c_uint8, C.c_uint16, C.c_uint32, C.c_float, C.c_int
    sz = C.c_size_t
    L.oracle_px_new.restype = C.POINTER(Px)
    L.oracle_px_new.argtypes = [f32, u16, u16, u8]
    L.oracle_px_delete.argtypes = [C.POINTER(Px)]
    L.oracle_px_time_mode.argtypes = [C.POINTER(Px), i32]
    L.oracle_px_node.restype = C.POINTER(Node)
    L.oracle_px_node.argtypes = [C.POINTER(Px), u32]
    L.oracle_get_d_from_intensity.restype = u8
    L.oracle_get_d_from_intensity.argtypes = [f32]
    L.oracle_px_pop_top_event.restype = Event
    L.oracle_px_pop_top_event.argtypes = [C.POINTER(Px), f32, i32, u32]
    L.oracle_px_pop_best_events.restype = i32
    L.oracle_px_pop_best_events.argtypes = [C.POINTER(Px), C.POINTER(Event), sz, i32, i32, u32, f32]
    L.oracle_px_set_d_for_continuous.restype = i32
    L.oracle_px_set_d_for_continuous.argtypes = [C.POINTER(Px), f32, u32, C.POINTER(Event)]
    L.oracle_px_integrate.argtypes = [C.POINTER(Px), f32, f32, i32, u32, u32, u8, u8, i32]
    L.oracle_is_feature.restype = i32
    L.oracle_is_feature.argtypes = [C.c_void_p, u16, u16, u8, u16, u16, u8]
    L.oracle_video_update_detect_features.restype = None
    L.oracle_video_update_detect_features.argtypes = [vp, i32, i32]
    L.oracle_video_new_features.restype = C.c_size_t
    L.oracle_video_new_features.argtypes = [vp, C.c_void_p, C.c_size_t]
    L.oracle_video_feature_mask.restype = C.POINTER(C.c_uint8)
    L.oracle_video_feature_mask.argtypes = [vp]
    i64 = C.c_int64
    L.oracle_framer_new.restype = vp
    L.oracle_framer_new.argtypes = [u16, u16, u8, u32, u8, i32, u32, u32, u32, f32, i32, u32, i64]
    L.oracle_framer_delete.argtypes = [vp]
    L.oracle_framer_ingest_event.restype = i32
    L.oracle_framer_ingest_event.argtypes = [vp, Event]
    L.oracle_framer_ingest_events_events.restype = i32
    L.oracle_framer_ingest_events_events.argtypes = [vp, C.c_void_p, C.c_void_p, u32]
    L.oracle_framer_write_multi_frame_bytes.restype = i32
    L.oracle_framer_write_multi_frame_bytes.argtypes = [vp, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.oracle_framer_flush_frame_buffer.restype = i32
    L.oracle_framer_flush_frame_buffer.argtypes = [vp]
    L.oracle_framer_bad.restype = i32
    L.oracle_framer_bad.argtypes = [vp]
    L.oracle_framer_frames_written.restype = i64
    L.oracle_framer_frames_written.argtypes = [vp]
    L.oracle_framer_tpf.restype = u32
    L.oracle_framer_tpf.argtypes = [vp]
    L.oracle_handle_color.restype = None
    L.oracle_handle_color.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.oracle_raw_header.restype = C.c_size_t
    L.oracle_raw_header.argtypes = [C.c_void_p, i32, u8, C.c_uint16, C.c_uint16, u32, u32, u32, u8, u32, u32, u32]
    L.oracle_raw_encode.restype = C.c_size_t
    L.oracle_raw_encode.argtypes = [C.c_void_p, C.c_size_t, u8, C.c_void_p]
    L.oracle_raw_eof.restype = C.c_size_t
    L.oracle_raw_eof.argtypes = [C.c_void_p]
    L.oracle_get_frame_value_u8.restype = u8
    L.oracle_get_frame_value_u8.argtypes = [u8, u32, C.c_double, f32, u32, i32, u32, u32]
    L.oracle_log2_raw.restype = f32
    L.oracle_log2_raw.argtypes = [f32]
    L.oracle_video_new.restype = vp
    L.oracle_video_new.argtypes = [u16, u16, u8, i32]
    L.oracle_video_delete.argtypes = [vp]
    L.oracle_video_chunk_rows.argtypes = [vp, u32]
    L.oracle_video_time_parameters.restype = i32
    L.oracle_video_time_parameters.argtypes = [vp, u32, u32, u32, i32]
    L.oracle_video_write_out.argtypes = [vp, i32, i32]
    L.oracle_video_update_crf.argtypes = [vp, u8]
    L.oracle_video_update_quality_manual.argtypes = [vp, u8, u8, u32, u8, f32]
    L.oracle_video_set_crf_parameters.argtypes = [vp, C.POINTER(CrfParameters)]
    L.oracle_video_update_delta_t_max.argtypes = [vp, u32]
    L.oracle_video_c_thresh_pos.argtypes = [vp, u8]
    L.oracle_video_set_c_thresh_rect.argtypes = [vp, u16, u16, u16, u16, u8]
    L.oracle_video_set_view_mode.argtypes = [vp, i32]
    L.oracle_video_set_in_interval_count.argtypes = [vp, u32]
    L.oracle_video_in_interval_count.restype = u32
    L.oracle_video_in_interval_count.argtypes = [vp]
    L.oracle_video_n_chunks.restype = u32
    L.oracle_video_n_chunks.argtypes = [vp]
    L.oracle_crf_parameters.argtypes = [u8, u16, u16, C.POINTER(CrfParameters)]
    L.oracle_video_integrate_matrix.restype = sz
    L.oracle_video_integrate_matrix.argtypes = [vp, vp, f32, i32]
    L.oracle_video_chunk_counts.argtypes = [vp, vp]
    L.oracle_video_copy_events.restype = sz
    L.oracle_video_copy_events.argtypes = [vp, vp, sz]
    L.oracle_video_running_intensities.restype = C.POINTER(C.c_uint8)
    L.oracle_video_running_intensities.argtypes = [vp]
    L.oracle_video_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(u32), C.POINTER(u32)]
    L.oracle_video_px.restype = C.POINTER(Px)
    L.oracle_video_px.argtypes = [vp, sz]
    L.oracle_max_threads.restype = i32
    L.oracle_num_procs.restype = i32
    L.oracle_synth_frame.restype = None
    L.oracle_synth_frame.argtypes = [i32, C.c_uint64, u32, u32, u32, C.c_uint64, C.c_uint64, vp, i32]
    _lib = L
    return L


class PixelArena:
    """One pixel, mirroring the reference's PixelArena test surface (event_pixel_tree.rs:534-1259)."""

    def __init__(self, start_intensity: float, x: int = 0, y: int = 0, c: int = C_NONE):
        self._L = lib()
        self._p = self._L.oracle_px_new(start_intensity, x, y, c)

    def __del__(self):
        try:
            self._L.oracle_px_delete(self._p)
        except Exception:
            pass

    def time_mode(self, mode):
        self._L.oracle_px_time_mode(self._p, -1 if mode is None else mode)

    def integrate(self, intensity, time, mode, dtm, ref_time, c_thresh_max, c_increase_velocity, multi_mode):
        self._L.oracle_px_integrate(self._p, intensity, time, mode, dtm, ref_time, c_thresh_max, c_increase_velocity, multi_mode)

    def pop_top_event(self, next_intensity, mode, ref_time):
        e = self._L.oracle_px_pop_top_event(self._p, next_intensity, mode, ref_time)
        return (e.d, e.t)

    def pop_best_events(self, mode, multi_mode, ref_time, intensity):
        buf = (Event * 64)()
        n = self._L.oracle_px_pop_best_events(self._p, buf, 64, mode, multi_mode, ref_time, intensity)
        assert n >= 0
        return [(buf[i].d, buf[i].t) for i in range(n)]

    def set_d_for_continuous(self, next_intensity, ref_time):
        e = Event()
        if self._L.oracle_px_set_d_for_continuous(self._p, next_intensity, ref_time, C.byref(e)):
            return (e.d, e.t)
        return None

    def node(self, idx) -> Node:
        return self._L.oracle_px_node(self._p, idx).contents

    @property
    def length(self):
        return self._p.contents.length

    @property
    def need_to_pop_top(self):
        return bool(self._p.contents.need_to_pop_top)

    @property
    def raw(self) -> Px:
        return self._p.contents


class Video:
    """Mirrors the transcode-state part of the reference's Video<W> (video.rs:322-345)."""

    def __init__(self, width, height, channels, pixel_tree_mode=MODE_FRAME_PERFECT):
        self._L = lib()
        self.w, self.h, self.c = width, height, channels
        self._v = self._L.oracle_video_new(width, height, channels, pixel_tree_mode)
        if not self._v:
            raise ValueError("invalid plane")

    def __del__(self):
        try:
            self._L.oracle_video_delete(self._v)
        except Exception:
            pass

    def chunk_rows(self, n):
        self._L.oracle_video_chunk_rows(self._v, n)
        return self

    def time_parameters(self, tps, ref_time, delta_t_max, time_mode=None):
        return bool(self._L.oracle_video_time_parameters(self._v, tps, ref_time, delta_t_max, -1 if time_mode is None else time_mode))

    def write_out(self, time_mode=None, pixel_multi_mode=None):
        self._L.oracle_video_write_out(self._v, -1 if time_mode is None else time_mode, -1 if pixel_multi_mode is None else pixel_multi_mode)

    def update_crf(self, crf):
        self._L.oracle_video_update_crf(self._v, crf)

    def update_quality_manual(self, c_base, c_max, dtm_mult, velocity, radius=0.0):
        self._L.oracle_video_update_quality_manual(self._v, c_base, c_max, dtm_mult, velocity, radius)

    def set_crf_parameters(self, c_base, c_max, velocity, radius=0):
        p = CrfParameters(c_base, c_max, velocity, 0, radius, 0)
        self._L.oracle_video_set_crf_parameters(self._v, C.byref(p))

    def update_delta_t_max(self, dtm):
        self._L.oracle_video_update_delta_t_max(self._v, dtm)

    def c_thresh_pos(self, c):
        self._L.oracle_video_c_thresh_pos(self._v, c)

    def set_c_thresh_rect(self, x0, y0, x1, y1, value):
        self._L.oracle_video_set_c_thresh_rect(self._v, x0, y0, x1, y1, value)

    def set_view_mode(self, m):
        self._L.oracle_video_set_view_mode(self._v, m)

    def set_in_interval_count(self, n):
        self._L.oracle_video_set_in_interval_count(self._v, n)

    @property
    def in_interval_count(self):
        return self._L.oracle_video_in_interval_count(self._v)

    @property
    def n_chunks(self):
        return self._L.oracle_video_n_chunks(self._v)

    def integrate_matrix(self, frame: np.ndarray, time_spanned: float, n_threads: int = 1):
        """Returns (events[EVENT_DTYPE], chunk_counts[u32])."""
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        assert frame.size == self.w * self.h * self.c
        total = self._L.oracle_video_integrate_matrix(self._v, frame.ctypes.data, time_spanned, n_threads)
        ev = np.empty(total, dtype=EVENT_DTYPE)
        got = self._L.oracle_video_copy_events(self._v, ev.ctypes.data, total)
        assert got == total
        counts = np.empty(self.n_chunks, dtype=np.uint32)
        self._L.oracle_video_chunk_counts(self._v, counts.ctypes.data)
        return ev, counts

    def integrate_matrix_count_only(self, frame: np.ndarray, time_spanned: float, n_threads: int = 1) -> int:
        """Same work, events stay in the per-chunk vectors (the timed CPU-baseline form)."""
        return self._L.oracle_video_integrate_matrix(self._v, frame.ctypes.data, time_spanned, n_threads)

    def update_detect_features(self, detect_features: bool, feature_rate_adjustment: bool = False):
        """video.rs:825-837"""
        self._L.oracle_video_update_detect_features(self._v, int(detect_features), int(feature_rate_adjustment))

    def new_features(self) -> np.ndarray:
        """[x, y] of the features newly inserted by the last integrate_matrix, sorted (the reference keeps a HashSet)."""
        n = self._L.oracle_video_new_features(self._v, None, 0)
        out = np.empty((n, 2), dtype=np.uint16)
        if n:
            self._L.oracle_video_new_features(self._v, out.ctypes.data, n)
        return out[np.lexsort((out[:, 0], out[:, 1]))] if n else out

    def feature_mask(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.oracle_video_feature_mask(self._v), shape=(self.h * self.w,)).reshape(self.h, self.w).copy()

    def running_intensities(self) -> np.ndarray:
        p = self._L.oracle_video_running_intensities(self._v)
        n = self.w * self.h * self.c
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.h, self.w, self.c).copy()

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        m, e = C.c_uint32(), C.c_uint32()
        self._L.oracle_video_stats(self._v, C.byref(a), C.byref(b), C.byref(m), C.byref(e))
        return {"live_nodes_entry": a.value, "live_nodes_exit": b.value, "max_live_nodes": m.value, "max_px_events": e.value}

    def px(self, index) -> Px:
        return self._L.oracle_video_px(self._v, index).contents


def crf_parameters(crf, w, h) -> CrfParameters:
    out = CrfParameters()
    lib().oracle_crf_parameters(crf, w, h, C.byref(out))
    return out


def max_threads() -> int:
    return lib().oracle_max_threads()


def host_threads() -> int:
    """The host cores this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, lib().oracle_num_procs())


def synth_frame(kind: int, seed: int, f: int, w: int, h: int, c: int, row0: int = 0, full_w: int = None, n_threads: int = 0) -> np.ndarray:
    """One (h, w, c) u8 synthetic frame — rows row0 .. row0+h of the plane — identical to tests/synth.py and the device generator."""
    out = np.empty((h, w, c), dtype=np.uint8)
    lib().oracle_synth_frame(kind, seed & 0xFFFFFFFFFFFFFFFF, f, w, c, row0 * w * c, out.size, out.ctypes.data, n_threads or host_threads())
    return out


def raw_header(width, height, channels, tps, ref_interval, delta_t_max, version=3, source_camera=0, time_mode=TIME_ABSOLUTE_T,
               adu_interval=0, compressed=False) -> bytes:
    """EventStreamHeader + extensions (codec/header.rs, encoder.rs:170-229)."""
    buf = (C.c_uint8 * 64)()
    n = lib().oracle_raw_header(buf, int(compressed), version, width, height, tps, ref_interval, delta_t_max, channels,
                                source_camera, time_mode, adu_interval)
    return bytes(buf[:n])


def raw_encode(events: np.ndarray, channels: int) -> bytes:
    """RawOutput::ingest_event over an event array (raw/stream.rs:100-120)."""
    events = np.ascontiguousarray(events)
    out = np.empty(len(events) * 11 + 1, dtype=np.uint8)
    n = lib().oracle_raw_encode(events.ctypes.data, len(events), channels, out.ctypes.data)
    return out[:n].tobytes()


def raw_eof() -> bytes:
    buf = (C.c_uint8 * 11)()
    lib().oracle_raw_eof(buf)
    return bytes(buf)


def handle_color(frame: np.ndarray) -> np.ndarray:
    """(H, W, 3) u8 -> (H, W, 1) u8 gray, utils/cv.rs:215-232."""
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    assert frame.shape[-1] == 3
    out = np.empty(frame.shape[:-1] + (1,), dtype=np.uint8)
    lib().oracle_handle_color(frame.ctypes.data, out.size, out.ctypes.data)
    return out


def is_feature_map(img: np.ndarray) -> np.ndarray:
    """oracle_is_feature at every pixel of an (H, W, C) u8 image -> (H, W) bool."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    L = lib()
    out = np.zeros((h, w), dtype=bool)
    for y in range(h):
        for x in range(w):
            out[y, x] = bool(L.oracle_is_feature(img.ctypes.data, w, h, c, x, y, 0))
    return out


class Framer:
    """FrameSequence<u8>, FramerMode::INSTANTANEOUS (framer/driver.rs), built like the reference's FramerBuilder chain."""

    def __init__(self, width, height, channels, chunk_rows, codec_version, time_mode, tps, ref_interval, delta_t_max,
                 output_fps=None, view_mode=VIEW_INTENSITY, source_camera=0, buffer_limit=None):
        self._L = lib()
        self.w, self.h, self.c = width, height, channels
        self._f = self._L.oracle_framer_new(width, height, channels, chunk_rows, codec_version, time_mode, tps, ref_interval,
                                            delta_t_max, 0.0 if output_fps is None else output_fps, view_mode, source_camera,
                                            -1 if buffer_limit is None else buffer_limit)
        assert self._f

    def __del__(self):
        try:
            self._L.oracle_framer_delete(self._f)
        except Exception:
            pass

    @property
    def tpf(self):
        return self._L.oracle_framer_tpf(self._f)

    @property
    def frames_written(self):
        return self._L.oracle_framer_frames_written(self._f)

    @property
    def bad(self):
        return bool(self._L.oracle_framer_bad(self._f))

    def ingest_event(self, x, y, c, d, t) -> bool:
        return bool(self._L.oracle_framer_ingest_event(self._f, Event(x, y, c, d, 0, t)))

    def ingest_events_events(self, events: np.ndarray, chunk_counts: np.ndarray) -> bool:
        events = np.ascontiguousarray(events)
        cc = np.ascontiguousarray(chunk_counts, dtype=np.uint32)
        return bool(self._L.oracle_framer_ingest_events_events(self._f, events.ctypes.data, cc.ctypes.data, len(cc)))

    def write_multi_frame_bytes(self, max_frames=1024) -> np.ndarray:
        """Pops every filled frame; returns them as (n, H, W, C) u8."""
        fb = self.w * self.h * self.c
        out = np.empty(max_frames * fb, dtype=np.uint8)
        nb = C.c_size_t()
        n = self._L.oracle_framer_write_multi_frame_bytes(self._f, out.ctypes.data, out.size, C.byref(nb))
        if n < 0:
            raise RuntimeError("write_multi_frame_bytes: reference error path (bad fill count / no frame) or buffer too small")
        return out[: nb.value].reshape(n, self.h, self.w, self.c).copy()

    def flush_frame_buffer(self) -> bool:
        return bool(self._L.oracle_framer_flush_frame_buffer(self._f))
