#!/bin/bash
# Kernel time against tiles per CTA: planes of W = 1984 (one 62-row tile per image row), H = rows.
set -u
for h in 592 1184 1776 2960 3256 3552 5920 11840; do
  echo "== H=$h ($(python -c "print($h/592)") tiles per CTA)"
  timeout 300 python tools/profile_run.py --reps 2 --w 1984 --h $h --c 1 --kind 1 --frames 40 2>&1 | tail -1
done
