#!/usr/bin/env python
"""Shares of executed instructions / stall samples per named line range.
usage: python tools/ncu_phases.py cs.csv file:a-b:name ..."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, fname, data = None, "", []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iI, iW = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        data.append((fname, int(r[0]), int(r[iI]), int(r[iW])))
    except ValueError:
        pass
tot, tots = sum(d[2] for d in data), sum(d[3] for d in data)
print(f"total warp-inst {tot}  stall samples {tots}")
seen = set()
for spec in sys.argv[2:]:
    f, rng, name = spec.split(":")
    a, b = map(int, rng.split("-"))
    sel = [d for d in data if d[0].startswith(f) and a <= d[1] <= b]
    seen.update((d[0], d[1]) for d in sel)
    print(f"{name:28} inst {sum(d[2] for d in sel) / tot * 100:5.1f}%  stall {sum(d[3] for d in sel) / tots * 100:5.1f}%")
rest = [d for d in data if (d[0], d[1]) not in seen]
print(f"{'(other)':28} inst {sum(d[2] for d in rest) / tot * 100:5.1f}%  stall {sum(d[3] for d in rest) / tots * 100:5.1f}%")
