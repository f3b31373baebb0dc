"""ctypes binding of libadder_b200.so (C ABI: include/adder_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("ADDER_B200_SO") or os.path.join(_HERE, "libadder_b200.so")  # override: A/B builds of the kernel
_CSRC = os.path.join(_HERE, "csrc")

# adder_event_t (12 bytes, little-endian)
EVENT_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("c", "u1"), ("d", "u1"), ("reserved", "<u2"), ("t", "<u4")])
assert EVENT_DTYPE.itemsize == 12

MODE_FRAME_PERFECT, MODE_CONTINUOUS = 0, 1
MULTI_NORMAL, MULTI_COLLAPSE = 0, 1
TIME_DELTA_T, TIME_ABSOLUTE_T, TIME_MIXED = 0, 1, 2
VIEW_INTENSITY, VIEW_D, VIEW_DELTA_T, VIEW_SAE = 0, 1, 2, 3

OK, ERR_BAD_PARAMS, ERR_NO_DEVICE, ERR_CUDA, ERR_CAPACITY, ERR_ARENA_DEPTH, ERR_UNSUPPORTED, ERR_INTERNAL, ERR_NOMEM = range(9)
_NAMES = ["OK", "BAD_PARAMS", "NO_DEVICE", "CUDA", "CAPACITY", "ARENA_DEPTH", "UNSUPPORTED", "INTERNAL", "NOMEM"]

# every symbol include/adder_b200.h declares (tests check the built library exports all of them)
SYMBOLS = [
    "adder_b200_abi_version", "adder_b200_last_error", "adder_b200_device_count", "adder_b200_crf_parameters",
    "adder_b200_video_create", "adder_b200_video_destroy", "adder_b200_video_chunk_rows",
    "adder_b200_video_time_parameters", "adder_b200_video_write_out", "adder_b200_video_update_crf",
    "adder_b200_video_update_quality_manual", "adder_b200_video_set_crf_parameters", "adder_b200_video_update_delta_t_max", "adder_b200_video_c_thresh_pos",
    "adder_b200_video_set_c_thresh_rect", "adder_b200_video_set_view_mode", "adder_b200_video_set_in_interval_count",
    "adder_b200_video_set_row_offset", "adder_b200_video_set_counting", "adder_b200_video_read_counters",
    "adder_b200_video_get_info", "adder_b200_video_integrate_matrix", "adder_b200_video_fetch_events",
    "adder_b200_video_running_intensities", "adder_b200_video_integrate_frames_device", "adder_b200_video_sync",
    "adder_b200_video_stream", "adder_b200_video_launch_count", "adder_b200_video_events_emitted",
    "adder_b200_video_integrate_frames_host", "adder_b200_video_reset_state", "adder_b200_video_read_px",
    "adder_b200_host_alloc", "adder_b200_host_free", "adder_b200_device_alloc", "adder_b200_device_free",
    "adder_b200_copy_to_device", "adder_b200_copy_to_host", "adder_b200_video_timer_start",
    "adder_b200_video_timer_stop", "adder_b200_synth_frames",
    "adder_b200_video_raw_header", "adder_b200_raw_eof", "adder_b200_video_raw_event_size",
    "adder_b200_video_raw_encode_device", "adder_b200_video_integrate_frames_host_raw",
    "adder_b200_video_set_source_channels", "adder_b200_video_input_frame",
    "adder_b200_video_update_detect_features", "adder_b200_video_new_features", "adder_b200_video_feature_mask",
    "adder_b200_framer_create", "adder_b200_framer_destroy", "adder_b200_framer_ingest_events_device",
    "adder_b200_framer_ingest_events_host", "adder_b200_framer_write_multi_frame_bytes",
    "adder_b200_framer_flush_frame_buffer", "adder_b200_framer_state",
    "adder_b200_comm_create", "adder_b200_comm_export", "adder_b200_comm_open", "adder_b200_comm_attach", "adder_b200_comm_destroy",
    "adder_b200_comm_push_frames", "adder_b200_comm_wait_frames", "adder_b200_comm_frame", "adder_b200_comm_release_frames",
    "adder_b200_comm_sync", "adder_b200_comm_stream",
    "adder_b200_video_integrate_frames_host_compact", "adder_b200_compact_frame_bytes", "adder_b200_expand_compact",
    "adder_b200_framer_ingest_events_device_async", "adder_b200_framer_frame_ready",
]
COMM_BLOB_BYTES = 256


class AdderError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"adder_b200: {_NAMES[code] if 0 <= code < len(_NAMES) else code}: {msg}")
        self.code = code


class CrfParameters(C.Structure):
    _fields_ = [("c_thresh_baseline", C.c_uint8), ("c_thresh_max", C.c_uint8), ("c_increase_velocity", C.c_uint8),
                ("reserved", C.c_uint8), ("feature_c_radius", C.c_uint16), ("reserved2", C.c_uint16)]


class VideoInfo(C.Structure):
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("channels", C.c_uint8), ("pixel_tree_mode", C.c_uint8),
                ("pixel_multi_mode", C.c_uint8), ("time_mode", C.c_uint8), ("view_mode", C.c_uint8),
                ("state_form", C.c_uint8), ("reserved", C.c_uint8 * 2), ("chunk_rows", C.c_uint32), ("n_chunks", C.c_uint32),
                ("in_interval_count", C.c_uint32), ("tps", C.c_uint32), ("ref_time", C.c_uint32),
                ("delta_t_max", C.c_uint32), ("crf", CrfParameters), ("max_depth", C.c_uint32), ("device", C.c_uint32),
                ("state_bytes", C.c_uint64), ("events_capacity", C.c_uint64)]


class PxNode(C.Structure):
    _fields_ = [("integration", C.c_float), ("delta_t", C.c_float), ("best_delta_t", C.c_float), ("d", C.c_uint8),
                ("best_d", C.c_uint8), ("has_best", C.c_uint8), ("reserved", C.c_uint8)]


class PxState(C.Structure):
    _fields_ = [("last_fired_t", C.c_float), ("running_t", C.c_float), ("base_val", C.c_uint8), ("c_thresh", C.c_uint8),
                ("c_increase_counter", C.c_uint8), ("length", C.c_uint8), ("dtm_reached", C.c_uint8),
                ("popped_dtm", C.c_uint8), ("time_mode", C.c_uint8), ("reserved", C.c_uint8), ("nodes", PxNode * 31)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libadder_b200.so for sm_100a with the committed recipe (csrc/Makefile)."""
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".h")) or f == "Makefile"]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "adder_b200.h"))
    stale = not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps)
    if (force or stale) and not os.environ.get("ADDER_B200_SO"):
        r = subprocess.run(["make", "-C", _CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("building libadder_b200.so failed")
    return _SO


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Fails loudly when it is missing: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the product has no CPU or PyTorch fallback)")
    L = C.CDLL(_SO)
    vp, u8, u16, u32, u64, f32, i32, sz = (C.c_void_p, C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_float,
                                           C.c_int, C.c_size_t)
    P = C.POINTER
    sig = {
        "adder_b200_abi_version": (i32, []),
        "adder_b200_last_error": (C.c_char_p, []),
        "adder_b200_device_count": (i32, []),
        "adder_b200_crf_parameters": (i32, [u8, u16, u16, P(CrfParameters)]),
        "adder_b200_video_create": (i32, [u16, u16, u8, i32, i32, u32, P(vp)]),
        "adder_b200_video_destroy": (None, [vp]),
        "adder_b200_video_chunk_rows": (i32, [vp, u32]),
        "adder_b200_video_time_parameters": (i32, [vp, u32, u32, u32, i32, P(i32)]),
        "adder_b200_video_write_out": (i32, [vp, i32, i32]),
        "adder_b200_video_update_crf": (i32, [vp, u8]),
        "adder_b200_video_update_quality_manual": (i32, [vp, u8, u8, u32, u8, f32]),
        "adder_b200_video_set_crf_parameters": (i32, [vp, P(CrfParameters)]),
        "adder_b200_video_update_delta_t_max": (i32, [vp, u32]),
        "adder_b200_video_c_thresh_pos": (i32, [vp, u8]),
        "adder_b200_video_set_c_thresh_rect": (i32, [vp, u16, u16, u16, u16, u8]),
        "adder_b200_video_set_view_mode": (i32, [vp, i32]),
        "adder_b200_video_set_in_interval_count": (i32, [vp, u32]),
        "adder_b200_video_set_row_offset": (i32, [vp, u16]),
        "adder_b200_video_set_counting": (i32, [vp, i32]),
        "adder_b200_video_read_counters": (i32, [vp, P(u64)]),
        "adder_b200_video_get_info": (i32, [vp, P(VideoInfo)]),
        "adder_b200_video_integrate_matrix": (i32, [vp, vp, sz, f32, vp, sz, vp, P(u64)]),
        "adder_b200_video_fetch_events": (i32, [vp, vp, sz, vp, P(u64)]),
        "adder_b200_video_running_intensities": (i32, [vp, vp]),
        "adder_b200_video_integrate_frames_device": (i32, [vp, vp, sz, u32, f32, vp, sz, vp]),
        "adder_b200_video_sync": (i32, [vp]),
        "adder_b200_video_stream": (vp, [vp]),
        "adder_b200_video_launch_count": (u64, [vp]),
        "adder_b200_video_events_emitted": (i32, [vp, P(u64)]),
        "adder_b200_video_integrate_frames_host": (i32, [vp, vp, sz, u32, f32, vp, sz, vp, vp, P(u64), P(u32)]),
        "adder_b200_video_reset_state": (i32, [vp]),
        "adder_b200_video_read_px": (i32, [vp, sz, P(PxState)]),
        "adder_b200_host_alloc": (i32, [sz, P(vp)]),
        "adder_b200_host_free": (i32, [vp]),
        "adder_b200_device_alloc": (i32, [vp, sz, P(vp)]),
        "adder_b200_device_free": (i32, [vp, vp]),
        "adder_b200_copy_to_device": (i32, [vp, vp, vp, sz]),
        "adder_b200_copy_to_host": (i32, [vp, vp, vp, sz]),
        "adder_b200_video_timer_start": (i32, [vp]),
        "adder_b200_video_timer_stop": (i32, [vp, P(f32)]),
        "adder_b200_synth_frames": (i32, [vp, vp, sz, u32, u32, i32, u64]),
        "adder_b200_video_raw_header": (i32, [vp, u8, u32, u32, vp, sz, P(sz)]),
        "adder_b200_raw_eof": (i32, [vp, sz, P(sz)]),
        "adder_b200_video_raw_event_size": (i32, [vp]),
        "adder_b200_video_raw_encode_device": (i32, [vp, vp, vp, u64, vp]),
        "adder_b200_video_integrate_frames_host_raw": (i32, [vp, vp, sz, u32, f32, vp, sz, vp, vp, P(u64), P(u32)]),
        "adder_b200_video_set_source_channels": (i32, [vp, u8]),
        "adder_b200_video_input_frame": (i32, [vp, vp]),
        "adder_b200_video_update_detect_features": (i32, [vp, i32, i32]),
        "adder_b200_video_new_features": (i32, [vp, vp, sz, P(u32)]),
        "adder_b200_video_feature_mask": (i32, [vp, vp]),
        "adder_b200_framer_create": (i32, [u16, u16, u8, u32, u8, i32, u32, u32, u32, f32, i32, u32, C.c_int64, u32, i32, P(vp)]),
        "adder_b200_framer_destroy": (None, [vp]),
        "adder_b200_framer_ingest_events_device": (i32, [vp, vp, vp, P(i32)]),
        "adder_b200_framer_ingest_events_host": (i32, [vp, vp, vp, P(i32)]),
        "adder_b200_framer_write_multi_frame_bytes": (i32, [vp, vp, u32, P(u32)]),
        "adder_b200_framer_flush_frame_buffer": (i32, [vp, P(i32)]),
        "adder_b200_framer_state": (i32, [vp, P(C.c_int64), P(u32)]),
        "adder_b200_comm_create": (i32, [vp, u32, u32, u32, sz, P(vp)]),
        "adder_b200_comm_export": (i32, [vp, vp, sz]),
        "adder_b200_comm_open": (i32, [vp, vp, sz, P(vp)]),
        "adder_b200_comm_attach": (i32, [vp, vp, P(vp)]),
        "adder_b200_comm_destroy": (None, [vp]),
        "adder_b200_comm_push_frames": (i32, [vp, u32, u32, vp, sz, vp, u32, u64]),
        "adder_b200_comm_wait_frames": (i32, [vp, u64, u32]),
        "adder_b200_comm_frame": (i32, [vp, u64, P(vp), P(vp)]),
        "adder_b200_comm_release_frames": (i32, [vp, u64]),
        "adder_b200_comm_sync": (i32, [vp]),
        "adder_b200_comm_stream": (vp, [vp]),
        "adder_b200_video_integrate_frames_host_compact": (i32, [vp, vp, sz, u32, f32, vp, sz, vp, vp, P(u64), P(u32)]),
        "adder_b200_compact_frame_bytes": (u64, [u64, u64]),
        "adder_b200_expand_compact": (i32, [u16, u16, u8, u16, vp, u64, vp, u32]),
        "adder_b200_framer_ingest_events_device_async": (i32, [vp, vp, vp]),
        "adder_b200_framer_frame_ready": (i32, [vp, P(i32)]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise AdderError(rc, lib().adder_b200_last_error().decode(errors="replace"))


def device_count() -> int:
    n = lib().adder_b200_device_count()
    return max(n, 0)


def raw_eof() -> bytes:
    """RawOutput::into_writer's EOF event (raw/stream.rs:79-92)."""
    buf = (C.c_uint8 * 16)()
    n = C.c_size_t()
    _check(lib().adder_b200_raw_eof(buf, 16, C.byref(n)))
    return bytes(buf[: n.value])


def compact_frame_bytes(n_px: int, n_events: int) -> int:
    return lib().adder_b200_compact_frame_bytes(n_px, n_events)


def expand_compact(width, rows, channels, row0, block: np.ndarray, n_events: int, out: np.ndarray = None, n_threads: int = 0) -> np.ndarray:
    """One frame's compact block -> its 12-byte records (host threads; no device involved)."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    if out is None:
        out = np.empty(n_events, dtype=EVENT_DTYPE)
    assert len(out) >= n_events
    _check(lib().adder_b200_expand_compact(width, rows, channels, row0, block.ctypes.data, n_events, out.ctypes.data, n_threads or (os.cpu_count() or 1)))
    return out[:n_events]


def crf_parameters(crf, w, h) -> CrfParameters:
    out = CrfParameters()
    _check(lib().adder_b200_crf_parameters(crf, w, h, C.byref(out)))
    return out


class _Pinned:
    def __init__(self, nbytes):
        self.L = lib()
        self.p = C.c_void_p()
        _check(self.L.adder_b200_host_alloc(nbytes, C.byref(self.p)))
        self.nbytes = nbytes

    def __del__(self):
        try:
            self.L.adder_b200_host_free(self.p)
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """A numpy array over page-locked host memory (adder_b200_host_alloc).  The allocation lives as
    long as any view of the array does (the ctypes buffer at the bottom of numpy's base chain owns it)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    owner = _Pinned(max(n, 1))
    buf = (C.c_uint8 * max(n, 1)).from_address(owner.p.value)
    buf._adder_owner = owner
    return np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)


class DeviceBuffer:
    """Device memory owned through the C ABI (no torch needed)."""

    def __init__(self, video: "Video", nbytes: int):
        self.video = video
        self.nbytes = nbytes
        self.p = C.c_void_p()
        _check(video.L.adder_b200_device_alloc(video.v, nbytes, C.byref(self.p)))

    @property
    def ptr(self):
        return self.p.value

    def to_host(self, dtype=np.uint8, nbytes=None, offset=0) -> np.ndarray:
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(nbytes, dtype=np.uint8)
        if nbytes:
            _check(self.video.L.adder_b200_copy_to_host(self.video.v, out.ctypes.data, self.p.value + offset, nbytes))
        return out.view(dtype)

    def to_host_into(self, out: np.ndarray, nbytes, offset=0):
        """Copy into an existing (e.g. page-locked) host array."""
        assert out.flags.c_contiguous and out.nbytes >= nbytes
        if nbytes:
            _check(self.video.L.adder_b200_copy_to_host(self.video.v, out.ctypes.data, self.p.value + offset, nbytes))

    def from_host(self, arr: np.ndarray, offset=0):
        arr = np.ascontiguousarray(arr)
        _check(self.video.L.adder_b200_copy_to_device(self.video.v, self.p.value + offset, arr.ctypes.data, arr.nbytes))

    def free(self):
        if self.p:
            self.video.L.adder_b200_device_free(self.video.v, self.p)
            self.p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Video:
    """Mirrors the transcode-state part of the reference's Video<W> (video.rs:322-345): same method
    names and argument meaning as the builder/setters of video.rs and the oracle's Video."""

    def __init__(self, width, height, channels, pixel_tree_mode=MODE_FRAME_PERFECT, device=0, max_depth=0):
        self.L = lib()
        self.w, self.h, self.c = width, height, channels
        self.src_c = channels  # channels of the frames handed in (set_source_channels)
        self.v = C.c_void_p()
        _check(self.L.adder_b200_video_create(width, height, channels, pixel_tree_mode, device, max_depth, C.byref(self.v)))

    def close(self):
        if getattr(self, "v", None):
            self.L.adder_b200_video_destroy(self.v)
            self.v = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- builder / setters ----
    def chunk_rows(self, n):
        _check(self.L.adder_b200_video_chunk_rows(self.v, n))
        return self

    def time_parameters(self, tps, ref_time, delta_t_max, time_mode=None) -> bool:
        applied = C.c_int()
        _check(self.L.adder_b200_video_time_parameters(self.v, tps, ref_time, delta_t_max, -1 if time_mode is None else time_mode, C.byref(applied)))
        return bool(applied.value)

    def write_out(self, time_mode=None, pixel_multi_mode=None):
        _check(self.L.adder_b200_video_write_out(self.v, -1 if time_mode is None else time_mode, -1 if pixel_multi_mode is None else pixel_multi_mode))

    def update_crf(self, crf):
        _check(self.L.adder_b200_video_update_crf(self.v, crf))

    def update_quality_manual(self, c_base, c_max, dtm_mult, velocity, radius=0.0):
        _check(self.L.adder_b200_video_update_quality_manual(self.v, c_base, c_max, dtm_mult, velocity, radius))

    def set_crf_parameters(self, c_base, c_max, velocity, radius=0):
        p = CrfParameters(c_base, c_max, velocity, 0, radius, 0)
        _check(self.L.adder_b200_video_set_crf_parameters(self.v, C.byref(p)))

    def update_delta_t_max(self, dtm):
        _check(self.L.adder_b200_video_update_delta_t_max(self.v, dtm))

    def c_thresh_pos(self, c):
        _check(self.L.adder_b200_video_c_thresh_pos(self.v, c))

    def set_c_thresh_rect(self, x0, y0, x1, y1, value):
        _check(self.L.adder_b200_video_set_c_thresh_rect(self.v, x0, y0, x1, y1, value))

    def set_view_mode(self, m):
        _check(self.L.adder_b200_video_set_view_mode(self.v, m))

    def set_in_interval_count(self, n):
        _check(self.L.adder_b200_video_set_in_interval_count(self.v, n))

    def set_source_channels(self, n):
        """3 on a one-channel video: frames come in as (H, W, 3) and handle_color (utils/cv.rs:215-232) runs on the device."""
        _check(self.L.adder_b200_video_set_source_channels(self.v, n))
        self.src_c = n if n else self.c

    def input_frame(self) -> np.ndarray:
        """The (gray) frame the last integrate call worked on (Framed.input_frame, framed.rs:129)."""
        out = np.empty((self.h, self.w, self.c), dtype=np.uint8)
        _check(self.L.adder_b200_video_input_frame(self.v, out.ctypes.data))
        return out

    def update_detect_features(self, detect_features: bool, feature_rate_adjustment: bool = False):
        """Video::update_detect_features, video.rs:825-837 (drawing / clustering flags are GUI-only)."""
        _check(self.L.adder_b200_video_update_detect_features(self.v, int(detect_features), int(feature_rate_adjustment)))

    def new_features(self) -> np.ndarray:
        """[x, y] of the features newly found by the last integrated frame, sorted (the reference keeps a HashSet)."""
        n = C.c_uint32()
        _check(self.L.adder_b200_video_new_features(self.v, None, 0, C.byref(n)))
        out = np.empty((n.value, 2), dtype=np.uint16)
        if n.value:
            _check(self.L.adder_b200_video_new_features(self.v, out.ctypes.data, n.value, C.byref(n)))
        return out[np.lexsort((out[:, 0], out[:, 1]))] if len(out) else out

    def feature_mask(self) -> np.ndarray:
        out = np.empty((self.h, self.w), dtype=np.uint8)
        _check(self.L.adder_b200_video_feature_mask(self.v, out.ctypes.data))
        return out

    def set_row_offset(self, row0):
        _check(self.L.adder_b200_video_set_row_offset(self.v, row0))

    def set_counting(self, on: bool):
        _check(self.L.adder_b200_video_set_counting(self.v, int(on)))

    def read_counters(self) -> dict:
        out = (C.c_uint64 * 6)()
        _check(self.L.adder_b200_video_read_counters(self.v, out))
        return dict(node_loads=out[0], node_stores=out[1], display_writes=out[2], events=out[3],
                    live_nodes_in=out[4], live_nodes_out=out[5])

    def reset_state(self):
        _check(self.L.adder_b200_video_reset_state(self.v))

    # ---- getters ----
    def info(self) -> VideoInfo:
        out = VideoInfo()
        _check(self.L.adder_b200_video_get_info(self.v, C.byref(out)))
        return out

    @property
    def state_form(self):
        """0: every level of a node stack holds its own values; 1: offset form (csrc/px_offset.cuh)."""
        return self.info().state_form

    @property
    def in_interval_count(self):
        return self.info().in_interval_count

    @property
    def n_chunks(self):
        return self.info().n_chunks

    @property
    def launch_count(self) -> int:
        return self.L.adder_b200_video_launch_count(self.v)

    def events_emitted(self) -> int:
        out = C.c_uint64()
        _check(self.L.adder_b200_video_events_emitted(self.v, C.byref(out)))
        return out.value

    def px(self, index) -> PxState:
        out = PxState()
        _check(self.L.adder_b200_video_read_px(self.v, index, C.byref(out)))
        return out

    def px_dict(self, index) -> dict:
        s = self.px(index)
        return dict(last_fired_t=s.last_fired_t, base_val=s.base_val, c_thresh=s.c_thresh,
                    c_increase_counter=s.c_increase_counter, length=s.length, popped_dtm=s.popped_dtm,
                    nodes=[dict(integration=n.integration, delta_t=n.delta_t, best_delta_t=n.best_delta_t, d=n.d,
                                best_d=n.best_d, has_best=n.has_best) for n in list(s.nodes)[:s.length]])

    # ---- the hot path ----
    def integrate_matrix(self, frame: np.ndarray, time_spanned: float, events_out: np.ndarray | None = None):
        """One frame, host buffers (Framed::consume's call, framed.rs:131).  Returns (events, chunk_counts)."""
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        assert frame.size == self.w * self.h * self.src_c
        counts = np.empty(self.n_chunks, dtype=np.uint32)
        n = C.c_uint64()
        if events_out is None:
            events_out = np.empty(self.w * self.h * self.c * 2, dtype=EVENT_DTYPE)
        rc = self.L.adder_b200_video_integrate_matrix(self.v, frame.ctypes.data, 0, time_spanned, events_out.ctypes.data,
                                                      len(events_out), counts.ctypes.data, C.byref(n))
        if rc == ERR_CAPACITY:  # nothing is lost: re-read into a buffer of the reported size
            events_out = np.empty(n.value, dtype=EVENT_DTYPE)
            rc = self.L.adder_b200_video_fetch_events(self.v, events_out.ctypes.data, len(events_out), counts.ctypes.data, C.byref(n))
        _check(rc)
        return events_out[: n.value], counts

    def integrate_frames_host(self, frames: np.ndarray, time_spanned: float, events_out: np.ndarray, partial: bool = False):
        """n frames, host buffers, copies and kernels pipelined.  Returns (events, frame_counts, chunk_counts).
        partial=True: a full events_out is not an error; returns (events, frame_counts, chunk_counts, frames_done) for the
        frames delivered — call again with frames[frames_done:] (the header's resume contract)."""
        assert frames.dtype == np.uint8 and frames.flags.c_contiguous
        nf = frames.shape[0]
        assert frames[0].size == self.w * self.h * self.src_c
        fc = np.zeros(nf, dtype=np.uint64)
        cc = np.zeros((nf, self.n_chunks), dtype=np.uint32)
        n, done = C.c_uint64(), C.c_uint32()
        rc = self.L.adder_b200_video_integrate_frames_host(self.v, frames.ctypes.data, frames[0].size, nf, time_spanned,
                                                           events_out.ctypes.data, len(events_out), fc.ctypes.data,
                                                           cc.ctypes.data, C.byref(n), C.byref(done))
        if rc == ERR_CAPACITY and partial:
            return events_out[: n.value], fc[: done.value], cc[: done.value], done.value
        _check(rc)
        return (events_out[: n.value], fc, cc, nf) if partial else (events_out[: n.value], fc, cc)

    # ---- raw .adder output ----
    def raw_header(self, version=3, source_camera=0, adu_interval=0) -> bytes:
        """EventStreamHeader + extensions for this plane and its time parameters (codec/header.rs, encoder.rs:170-229)."""
        buf = (C.c_uint8 * 64)()
        n = C.c_size_t()
        _check(self.L.adder_b200_video_raw_header(self.v, version, source_camera, adu_interval, buf, 64, C.byref(n)))
        return bytes(buf[: n.value])

    @property
    def raw_event_size(self) -> int:
        return self.L.adder_b200_video_raw_event_size(self.v)

    def raw_encode_device(self, d_events, d_n_events, n_events_max, d_out):
        """Records in HBM -> wire bytes in HBM (RawOutput::ingest_event, raw/stream.rs:100-120), on the handle's stream."""
        _check(self.L.adder_b200_video_raw_encode_device(self.v, d_events, d_n_events, n_events_max, d_out))

    def integrate_frames_host_raw(self, frames: np.ndarray, time_spanned: float, bytes_out: np.ndarray):
        """integrate_frames_host delivering the raw stream body.  Returns (bytes, frame_counts, chunk_counts)."""
        assert frames.dtype == np.uint8 and frames.flags.c_contiguous and bytes_out.dtype == np.uint8
        nf = frames.shape[0]
        assert frames[0].size == self.w * self.h * self.src_c
        fc = np.zeros(nf, dtype=np.uint64)
        cc = np.zeros((nf, self.n_chunks), dtype=np.uint32)
        n, done = C.c_uint64(), C.c_uint32()
        _check(self.L.adder_b200_video_integrate_frames_host_raw(self.v, frames.ctypes.data, frames[0].size, nf, time_spanned,
                                                                 bytes_out.ctypes.data, bytes_out.size, fc.ctypes.data,
                                                                 cc.ctypes.data, C.byref(n), C.byref(done)))
        return bytes_out[: n.value], fc, cc

    def integrate_frames_host_compact(self, frames: np.ndarray, time_spanned: float, bytes_out: np.ndarray):
        """integrate_frames_host delivering the compact form (include/adder_b200.h): per frame either count bytes + {d, t}
        or {index, d, t}.  Returns (bytes, frame_counts, chunk_counts); split with compact_frame_bytes, expand with expand_compact."""
        assert frames.dtype == np.uint8 and frames.flags.c_contiguous and bytes_out.dtype == np.uint8
        nf = frames.shape[0]
        assert frames[0].size == self.w * self.h * self.src_c
        fc = np.zeros(nf, dtype=np.uint64)
        cc = np.zeros((nf, self.n_chunks), dtype=np.uint32)
        n, done = C.c_uint64(), C.c_uint32()
        _check(self.L.adder_b200_video_integrate_frames_host_compact(self.v, frames.ctypes.data, frames[0].size, nf, time_spanned,
                                                                     bytes_out.ctypes.data, bytes_out.size, fc.ctypes.data,
                                                                     cc.ctypes.data, C.byref(n), C.byref(done)))
        return bytes_out[: n.value], fc, cc

    def running_intensities(self) -> np.ndarray:
        out = np.empty((self.h, self.w, self.c), dtype=np.uint8)
        _check(self.L.adder_b200_video_running_intensities(self.v, out.ctypes.data))
        return out

    # ---- device-resident form ----
    def device_alloc(self, nbytes) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def synth_frames(self, dbuf: DeviceBuffer, f0, n_frames, kind, seed, frame_stride=0, offset=0):
        _check(self.L.adder_b200_synth_frames(self.v, dbuf.ptr + offset, frame_stride, f0, n_frames, kind, seed))

    def integrate_frames_device(self, d_frames, frame_stride, n_frames, time_spanned, d_events, events_stride, d_chunk_offsets=None):
        _check(self.L.adder_b200_video_integrate_frames_device(self.v, d_frames, frame_stride, n_frames, time_spanned,
                                                               d_events, events_stride, d_chunk_offsets))

    def sync(self):
        _check(self.L.adder_b200_video_sync(self.v))

    def timer_start(self):
        _check(self.L.adder_b200_video_timer_start(self.v))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(self.L.adder_b200_video_timer_stop(self.v, C.byref(ms)))
        return ms.value


class Framer:
    """Mirrors the reference's FrameSequence<u8> in FramerMode::INSTANTANEOUS (framer/driver.rs) as built by
    FramerBuilder::new(plane, chunk_rows).codec_version(..).time_parameters(..).mode(INSTANTANEOUS).view_mode(..)
    .source(U8, source_camera).buffer_limit(..).finish()."""

    def __init__(self, width, height, channels, chunk_rows, codec_version, time_mode, tps, ref_interval, delta_t_max,
                 output_fps=None, view_mode=VIEW_INTENSITY, source_camera=0, buffer_limit=None, ring_frames=0, device=0):
        self.L = lib()
        self.w, self.h, self.c, self.chunk_rows = width, height, channels, chunk_rows
        self.n_chunks = (height + chunk_rows - 1) // chunk_rows
        self.f = C.c_void_p()
        _check(self.L.adder_b200_framer_create(width, height, channels, chunk_rows, codec_version, time_mode, tps, ref_interval,
                                               delta_t_max, 0.0 if output_fps is None else output_fps, view_mode, source_camera,
                                               -1 if buffer_limit is None else buffer_limit, ring_frames, device, C.byref(self.f)))

    def close(self):
        if getattr(self, "f", None):
            self.L.adder_b200_framer_destroy(self.f)
            self.f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _state(self):
        fw, tpf = C.c_int64(), C.c_uint32()
        _check(self.L.adder_b200_framer_state(self.f, C.byref(fw), C.byref(tpf)))
        return fw.value, tpf.value

    @property
    def frames_written(self):
        return self._state()[0]

    @property
    def tpf(self):
        return self._state()[1]

    def ingest_events_events(self, events: np.ndarray, chunk_counts: np.ndarray) -> bool:
        """Vec<Vec<Event>> as the concatenated records + per-chunk lengths (host memory)."""
        events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        cc = np.ascontiguousarray(chunk_counts, dtype=np.uint32)
        assert len(cc) == self.n_chunks and int(cc.sum()) == len(events)
        ready = C.c_int()
        _check(self.L.adder_b200_framer_ingest_events_host(self.f, events.ctypes.data if len(events) else None, cc.ctypes.data, C.byref(ready)))
        return bool(ready.value)

    def ingest_events_device(self, d_events, d_chunk_offsets) -> bool:
        """The transcoder's device-resident output (records + n_chunks+1 offsets in HBM)."""
        ready = C.c_int()
        _check(self.L.adder_b200_framer_ingest_events_device(self.f, d_events, d_chunk_offsets, C.byref(ready)))
        return bool(ready.value)

    def ingest_events_device_async(self, d_events, d_chunk_offsets):
        """ingest_events_device without the wait; ask frame_ready() later."""
        _check(self.L.adder_b200_framer_ingest_events_device_async(self.f, d_events, d_chunk_offsets))

    def frame_ready(self) -> bool:
        """is_frame_0_filled() as of the last ingest (the one synchronisation of the asynchronous form)."""
        ready = C.c_int()
        _check(self.L.adder_b200_framer_frame_ready(self.f, C.byref(ready)))
        return bool(ready.value)

    def ingest_event(self, x, y, c, d, t) -> bool:
        """Framer::ingest_event: one event (driver.rs:437-562)."""
        ev = np.zeros(1, dtype=EVENT_DTYPE)
        ev[0] = (x, y, c, d, 0, t)
        cc = np.zeros(self.n_chunks, dtype=np.uint32)
        if y // self.chunk_rows < self.n_chunks:
            cc[y // self.chunk_rows] = 1
        else:
            ev = ev[:0]
        return self.ingest_events_events(ev, cc)

    def write_multi_frame_bytes(self, max_frames=1024) -> np.ndarray:
        """FrameSequence::write_multi_frame_bytes (driver.rs:971-982): every finished frame, (n, H, W, C) u8.  The frames
        come through a small page-locked staging array, a few per call of the C entry point."""
        if getattr(self, "_stage", None) is None:
            self._stage = pinned_empty((4, self.h, self.w, self.c), np.uint8)
        got = []
        n = C.c_uint32()
        while len(got) < max_frames:
            want = min(len(self._stage), max_frames - len(got))
            _check(self.L.adder_b200_framer_write_multi_frame_bytes(self.f, self._stage.ctypes.data, want, C.byref(n)))
            got.extend(self._stage[k].copy() for k in range(n.value))
            if n.value < want:
                break
        if not got:
            return np.empty((0, self.h, self.w, self.c), dtype=np.uint8)
        return np.stack(got)

    def flush_frame_buffer(self) -> bool:
        ready = C.c_int()
        _check(self.L.adder_b200_framer_flush_frame_buffer(self.f, C.byref(ready)))
        return bool(ready.value)


class Exchange:
    """The event exchange between row bands (include/adder_b200.h, comm section): the consumer's ring of whole-frame
    buffers, or a band's mapping of it.  Build with Exchange.consumer(...), then .export() / Exchange.open(...) across
    processes or .attach(...) inside one."""

    def __init__(self, handle, video, owner):
        self.L = lib()
        self.c = handle
        self.video = video
        self.owner = owner

    @classmethod
    def consumer(cls, video: "Video", world: int, total_chunks: int, slots: int, out_stride: int) -> "Exchange":
        h = C.c_void_p()
        _check(video.L.adder_b200_comm_create(video.v, world, total_chunks, slots, out_stride, C.byref(h)))
        return cls(h, video, True)

    def export(self) -> bytes:
        buf = (C.c_uint8 * COMM_BLOB_BYTES)()
        _check(self.L.adder_b200_comm_export(self.c, buf, COMM_BLOB_BYTES))
        return bytes(buf)

    @classmethod
    def open(cls, video: "Video", blob: bytes) -> "Exchange":
        h = C.c_void_p()
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(video.L.adder_b200_comm_open(video.v, buf, len(blob), C.byref(h)))
        return cls(h, video, False)

    def attach(self, video: "Video") -> "Exchange":
        assert self.owner
        h = C.c_void_p()
        _check(self.L.adder_b200_comm_attach(video.v, self.c, C.byref(h)))
        return Exchange(h, video, False)

    def push_frames(self, band, chunk0, d_events, events_stride, d_chunk_offsets, n_frames, frame_seq0):
        _check(self.L.adder_b200_comm_push_frames(self.c, band, chunk0, d_events, events_stride, d_chunk_offsets, n_frames, frame_seq0))

    def wait_frames(self, frame_seq0, n_frames):
        _check(self.L.adder_b200_comm_wait_frames(self.c, frame_seq0, n_frames))

    def release_frames(self, upto_seq):
        _check(self.L.adder_b200_comm_release_frames(self.c, upto_seq))

    def frame_ptrs(self, frame_seq):
        e, o = C.c_void_p(), C.c_void_p()
        _check(self.L.adder_b200_comm_frame(self.c, frame_seq, C.byref(e), C.byref(o)))
        return e.value, o.value

    def read_frame(self, frame_seq, total_chunks):
        """Consumer, after wait_frames + sync: (events, chunk_offsets) of one whole frame as numpy arrays."""
        e, o = self.frame_ptrs(frame_seq)
        off = np.empty(total_chunks + 1, dtype=np.uint32)
        _check(self.L.adder_b200_copy_to_host(self.video.v, off.ctypes.data, o, off.nbytes))
        ev = np.empty(int(off[-1]), dtype=EVENT_DTYPE)
        if len(ev):
            _check(self.L.adder_b200_copy_to_host(self.video.v, ev.ctypes.data, e, ev.nbytes))
        return ev, off

    def sync(self):
        _check(self.L.adder_b200_comm_sync(self.c))

    def close(self):
        if getattr(self, "c", None):
            self.L.adder_b200_comm_destroy(self.c)
            self.c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
