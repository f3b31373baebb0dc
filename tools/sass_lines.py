#!/usr/bin/env python
"""Static SASS instruction count per source line for one kernel of a built library (no GPU needed).
usage: tools/sass_lines.py <lib.so> [kernel-mangled-substring] [file:first-last ...]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so = os.path.abspath(sys.argv[1])
pat = sys.argv[2] if len(sys.argv) > 2 else "integrate_frame_kernelILi8ELb0ELb1"
ranges = sys.argv[3:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and pat in l][0]
end = [i for i, l in enumerate(sass) if i > start and l.startswith("//--------------------- .text.")]
k = sass[start:end[0] if end else len(sass)]
cur, cnt, total, spills = None, collections.Counter(), 0, 0
for l in k:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m and "inlined" not in l:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l):
        total += 1
        cnt[cur] += 1
        spills += ("STL" in l) or ("LDL" in l)
print(f"kernel {pat}: {total} instructions, {spills} local-memory instructions")
for r in ranges:
    f, span = r.split(":")
    a, b = [int(x) for x in span.split("-")]
    t = sum(c for (ff, ln), c in cnt.items() if ff == f and a <= ln <= b)
    print(f"  {r}: {t}")
    if os.environ.get("VERBOSE"):
        src = None
        for root in ("adder_codec_rs_b200/csrc",):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
            if os.path.exists(p):
                src = open(p).read().split("\n")
        for (ff, ln), c in sorted(cnt.items()):
            if ff == f and a <= ln <= b:
                print(f"    {ln:4d} {c:3d}  {src[ln - 1].strip()[:110] if src else ''}")
