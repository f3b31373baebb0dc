"""CPU-side checks of the boundary: the library builds, loads, exports every symbol the header
declares, and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import adder_codec_rs_b200 as A
from adder_codec_rs_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    h = open(os.path.join(ROOT, "include", "adder_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(adder_b200_\w+)\s*\(", h)))


def test_library_builds_and_exports_every_header_symbol():
    B.build()
    L = C.CDLL(os.path.join(ROOT, "adder_codec_rs_b200", "libadder_b200.so"))
    names = _header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/adder_b200.h but not exported"
    assert sorted(B.SYMBOLS) == names, "binding.SYMBOLS out of sync with the header"


def test_abi_version_and_crf_table():
    L = A.lib()
    assert L.adder_b200_abi_version() == 1
    # rate_controller.rs:5-18 rows, and Crf::new's radius = denom * min(w,h)
    from oracle import oracle_py as O

    for crf in range(10):
        g, o = A.crf_parameters(crf, 1920, 1080), O.crf_parameters(crf, 1920, 1080)
        assert (g.c_thresh_baseline, g.c_thresh_max, g.c_increase_velocity, g.feature_c_radius) == (
            o.c_thresh_baseline, o.c_thresh_max, o.c_increase_velocity, o.feature_c_radius)
    p = A.crf_parameters(3, 640, 480)
    assert (p.c_thresh_baseline, p.c_thresh_max, p.c_increase_velocity) == (2, 7, 7)


def test_event_record_layout():
    assert A.EVENT_DTYPE.itemsize == 12
    assert [A.EVENT_DTYPE.fields[k][1] for k in ("x", "y", "c", "d", "reserved", "t")] == [0, 2, 4, 5, 6, 8]


@pytest.mark.skipif(A.device_count() > 0, reason="needs a GPU-less host")
def test_no_cpu_fallback():
    with pytest.raises(A.AdderError) as e:
        A.Video(16, 16, 1)
    assert e.value.code == B.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ (and tests/host_sim)."""
    pkg = os.path.join(ROOT, "adder_codec_rs_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dp, f), errors="replace").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"oracle_py", r"#include.*oracle", r"libadder_oracle", r"libpx_sim",
                            r"^\s*(from|import)\s+tests"):
                    assert not re.search(pat, text, flags=re.M), f"{f} reaches into test infrastructure ({pat})"


def _header_prototypes():
    h = open(os.path.join(ROOT, "include", "adder_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(adder_b200_\w+)\s*\(([^;{]*?)\)\s*;", h):
        args = " ".join(m.group(2).split())
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_bindings_cover_the_header():
    """rust/adder_b200-sys/src/lib.rs (source only: no rustc in this image) declares every entry point of the header
    with the same number of arguments, and its event record is the 12-byte layout."""
    src = open(os.path.join(ROOT, "rust", "adder_b200-sys", "src", "lib.rs")).read()
    rust = {}
    for m in re.finditer(r"pub fn (adder_b200_\w+)\(([^)]*)\)", src):
        args = m.group(2).strip()
        rust[m.group(1)] = 0 if not args else args.count(":")
    protos = _header_prototypes()
    assert set(rust) == set(protos), sorted(set(rust) ^ set(protos))
    for name, n in protos.items():
        assert rust[name] == n, f"{name}: header has {n} arguments, lib.rs {rust[name]}"
    ev = re.search(r"pub struct adder_event_t \{(.*?)\}", src, flags=re.S).group(1)
    fields = re.findall(r"pub (\w+): (\w+)", ev)
    assert fields == [("x", "u16"), ("y", "u16"), ("c", "u8"), ("d", "u8"), ("reserved", "u16"), ("t", "u32")]
