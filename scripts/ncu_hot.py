#!/usr/bin/env python
"""Compact view of an ncu source-page CSV: the SASS instructions holding the most stall samples, with their
neighbours' opcodes.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name K | python scripts/ncu_hot.py [top]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
S, I, SRC = ix["# Samples"], ix["Instructions Executed"], ix["Source"]
data = [r for r in rows[hi + 1:] if len(r) > max(S, I)]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


ts, ti = sum(num(r[S]) for r in data), sum(num(r[I]) for r in data)
print(f"{len(data)} SASS instructions, {ts} samples, {ti} warp instructions executed")
order = sorted(range(len(data)), key=lambda k: -num(data[k][S]))[:top]
for k in sorted(order):
    r = data[k]
    print(f"{k:6d} {100 * num(r[S]) / ts:5.1f}% smp {100 * num(r[I]) / ti:5.2f}% ins  {r[SRC][:100]}")
