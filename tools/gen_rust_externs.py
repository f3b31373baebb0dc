#!/usr/bin/env python
"""Rewrites the `extern "C" { ... }` block of rust/adder_b200-sys/src/lib.rs from include/adder_b200.h (one declaration per
entry point, in the header's order).  Run after adding an entry point to the header."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
h = open(os.path.join(ROOT, "include", "adder_b200.h")).read()
body = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
protos = re.findall(r"\n((?:int|void\*?|const char\*|uint64_t|void)\s+\*?adder_b200_\w+\s*\([^;]*?\))\s*;", body)
tmap = {"uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "size_t": "usize", "int": "c_int",
        "float": "c_float", "char": "c_char", "void": "c_void", "adder_event_t": "adder_event_t",
        "adder_crf_parameters_t": "adder_crf_parameters_t", "adder_b200_video": "adder_b200_video", "adder_b200_framer": "adder_b200_framer",
        "adder_b200_comm": "adder_b200_comm", "adder_b200_video_info_t": "adder_b200_video_info_t",
        "adder_b200_px_state_t": "adder_b200_px_state_t"}


def conv(t):
    t = t.strip()
    const = t.startswith("const ")
    if const:
        t = t[6:].strip()
    stars = t.count("*")
    r = tmap[t.replace("*", "").strip()]
    for _ in range(stars):
        r = ("*const " if const else "*mut ") + r
        const = False
    return r


out = []
for p in protos:
    p = " ".join(p.split())
    m = re.match(r"(.*?)\s*\*?(adder_b200_\w+)\s*\((.*)\)$", p)
    ret, name, args = m.group(1), m.group(2), m.group(3)
    if "*" in p[: p.index(name)] and not ret.endswith("*"):
        ret += "*"
    rargs = []
    if args.strip() != "void":
        for a in args.split(","):
            mm = re.match(r"(.*?)(\w+)(\[\d*\])?$", a.strip())
            ty, nm, arr = mm.group(1).strip(), mm.group(2), mm.group(3)
            if arr:
                ty += "*"
            if nm in ("ref", "type", "in", "out"):
                nm += "_"
            rargs.append(f"{nm}: {conv(ty)}")
    rret = "" if ret.strip() == "void" else " -> " + conv(ret)
    out.append(f"    pub fn {name}({', '.join(rargs)}){rret};")
path = os.path.join(ROOT, "rust", "adder_b200-sys", "src", "lib.rs")
src = open(path).read()
head = src[: src.index('extern "C" {')]
open(path, "w").write(head + 'extern "C" {\n' + "\n".join(out) + "\n}\n")
print(len(out), "entry points")
