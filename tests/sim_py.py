"""ctypes driver of tests/host_sim (the product's px_machine.cuh compiled for the host). TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle_py as O

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_sim")
_SO = os.path.join(_HERE, "libpx_sim.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_DEPS = [os.path.join(_HERE, "px_sim.cpp"),
         os.path.join(_ROOT, "adder_codec_rs_b200", "csrc", "px_machine.cuh"),
         os.path.join(_ROOT, "adder_codec_rs_b200", "csrc", "px_offset.cuh"),
         os.path.join(_ROOT, "adder_codec_rs_b200", "csrc", "state_layout.h"),
         os.path.join(_ROOT, "adder_codec_rs_b200", "csrc", "gray_math.h"),
         os.path.join(_ROOT, "adder_codec_rs_b200", "csrc", "raw_pack.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _DEPS):
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                            "-x", "c++", _DEPS[0], "-o", _SO], check=True)
        L = C.CDLL(_SO)
        u32, f32, vp, sz, i32 = C.c_uint32, C.c_float, C.c_void_p, C.c_size_t, C.c_int
        L.sim_new.restype = vp
        L.sim_new.argtypes = [u32, u32, u32, u32]
        L.sim_delete.argtypes = [vp]
        L.sim_reset_c.argtypes = [vp, u32, i32]
        L.sim_integrate.restype = sz
        L.sim_integrate.argtypes = [vp, vp, f32, u32, u32, u32, u32, i32, i32, i32, f32]
        L.sim_events.restype = vp
        L.sim_events.argtypes = [vp]
        L.sim_running.restype = C.POINTER(C.c_uint8)
        L.sim_running.argtypes = [vp]
        L.sim_force_display.argtypes = [vp]
        L.sim_set_fast_div_ulps.argtypes = [i32]
        L.sim_set_entry.argtypes = [i32]
        L.sim_gray_check.restype = C.c_uint64
        L.sim_gray_check.argtypes = [C.POINTER(C.c_uint64)]
        L.sim_gray_of.restype = u32
        L.sim_gray_of.argtypes = [u32, u32, u32]
        L.sim_raw_pack.argtypes = [vp, sz, u32, vp]
        L.sim_div_ref.restype = u32
        L.sim_div_ref.argtypes = [u32, u32]
        L.sim_frame_value_intensity.restype = u32
        L.sim_frame_value_intensity.argtypes = [u32, u32, u32]
        L.sim_form.restype = i32
        L.sim_form.argtypes = [vp]
        L.sim_rec_traffic.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.sim_err.restype = u32
        L.sim_err.argtypes = [vp]
        L.sim_px.argtypes = [vp, sz, C.POINTER(f32), C.POINTER(u32)]
        L.sim_node.argtypes = [vp, sz, u32, C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), C.POINTER(u32)]
        _lib = L
    return _lib


class SimVideo:
    def __init__(self, w, h, c, depth=12):
        self.L = lib()
        self.w, self.h, self.c, self.depth = w, h, c, depth
        self.v = self.L.sim_new(w, h, c, depth)
        self.ref, self.dtm, self.c_max, self.vel = 255, 7650, 7, 7
        self.collapse, self.abs_time, self.view = 1, 1, 0

    def __del__(self):
        try:
            self.L.sim_delete(self.v)
        except Exception:
            pass

    def force_display(self):
        """What the C ABI does after any setter that changes ref / dtm / view mode."""
        self.L.sim_force_display(self.v)

    def reset_c(self, c, reset_counter=True):
        self.L.sim_reset_c(self.v, c, int(reset_counter))

    def integrate(self, frame, time):
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        pdm = O.lib().oracle_log2_raw(255.0 * float(self.dtm // self.ref))
        n = self.L.sim_integrate(self.v, frame.ctypes.data, time, self.ref, self.dtm, self.c_max, self.vel, self.collapse,
                                 self.abs_time, self.view, pdm)
        p = self.L.sim_events(self.v)
        ev = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * 12,)).view(O.EVENT_DTYPE).copy() if n else np.empty(0, O.EVENT_DTYPE)
        return ev

    def running(self):
        n = self.w * self.h * self.c
        return np.ctypeslib.as_array(self.L.sim_running(self.v), shape=(n,)).reshape(self.h, self.w, self.c).copy()

    @property
    def form(self):
        """-1 undecided, 0 eager, 1 offset form (entry 4 only)."""
        return self.L.sim_form(self.v)

    def rec_traffic(self):
        out = (C.c_uint64 * 3)()
        self.L.sim_rec_traffic(self.v, out)
        return tuple(out)

    @property
    def err(self):
        return self.L.sim_err(self.v)

    def px(self, i):
        lf, y = C.c_float(), C.c_uint32()
        self.L.sim_px(self.v, i, C.byref(lf), C.byref(y))
        y = y.value
        st = dict(last_fired_t=lf.value, base_val=y & 0xFF, c_thresh=(y >> 8) & 0xFF, c_increase_counter=(y >> 16) & 0xFF,
                  length=(y >> 24) & 0x1F, dtm_reached=(y >> 29) & 1, popped_dtm=(y >> 30) & 1, nodes=[])
        for k in range(st["length"]):
            a, b, c_, w = C.c_float(), C.c_float(), C.c_float(), C.c_uint32()
            self.L.sim_node(self.v, i, k, C.byref(a), C.byref(b), C.byref(c_), C.byref(w))
            st["nodes"].append(dict(integration=a.value, delta_t=b.value, best_delta_t=c_.value, d=w.value & 0xFF,
                                    best_d=(w.value >> 8) & 0xFF, has_best=(w.value >> 16) & 1))
        return st
