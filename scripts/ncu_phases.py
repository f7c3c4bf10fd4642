#!/usr/bin/env python
"""Per-phase (between BAR.SYNC instructions) shares of stall samples and executed instructions of one kernel.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name K | python scripts/ncu_phases.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
S, I, SRC, T = ix["# Samples"], ix["Instructions Executed"], ix["Source"], ix["Thread Instructions Executed"]
data = [r for r in rows[hi + 1:] if len(r) > max(S, I)]
if len(data) % 2 == 0 and [r[SRC] for r in data[:len(data) // 2]] == [r[SRC] for r in data[len(data) // 2:]]:
    data = data[:len(data) // 2]      # ncu lists the function twice


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


ts, ti = sum(num(r[S]) for r in data), sum(num(r[I]) for r in data)
print(f"{len(data)} SASS instructions, {ts} samples, {ti} warp instructions")
seg = s = i = t = start = 0
ops = {}
for k, r in enumerate(data):
    s += num(r[S]); i += num(r[I]); t += num(r[T])
    op = (r[SRC].split() or ["?"])
    op = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    for key in ("LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "UBLKPF"):
        if op.startswith(key):
            ops[key] = ops.get(key, 0) + num(r[I])
    if "BAR.SYNC" in r[SRC] or k == len(data) - 1:
        print(f"phase {seg:2d} [{start:5d}-{k:5d}] {100 * s / ts:5.1f}% smp {100 * i / ti:5.1f}% ins  lanes/instr {t / max(i, 1):4.1f}  "
              + " ".join(f"{a}:{100 * b / ti:.1f}%" for a, b in ops.items()))
        seg += 1; s = i = t = 0; start = k + 1; ops = {}
