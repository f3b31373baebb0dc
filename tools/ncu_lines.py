#!/usr/bin/env python
"""Per-source-line shares of executed instructions and stall samples from an ncu report.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::1 > cs.csv; python tools/ncu_lines.py cs.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out, fname = [], ""
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed")
        iW = hdr.index("Warp Stall Sampling (All Samples)")
        names = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        n, s = int(r[iI]), int(r[iW])
    except ValueError:
        continue
    st = {}
    for i, h in names:
        try:
            st[h] = int(r[i])
        except ValueError:
            pass
    out.append((fname, int(r[0]), n, s, r[1], st))
tot = sum(o[2] for o in out)
tots = sum(o[3] for o in out)
print(f"total warp-inst {tot}, stall samples {tots}")
for f, ln, n, s, src, st in out:
    if n > tot * thr / 100 or s > tots * thr / 100:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        tops = " ".join(f"{k[6:]}:{v}" for k, v in top if v)
        print(f"{f[:14]:14}:{ln:<4} inst {n / tot * 100:5.1f}%  stall {s / tots * 100:5.1f}%  [{tops}]  {src.strip()[:100]}")
