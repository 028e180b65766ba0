#!/usr/bin/env python
"""Top source lines / SASS instructions by stall samples from `ncu -i rep --page source --csv --print-source {cuda|sass}`.
usage: src_lines.py dump.csv [top] [context]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = None
for i, r in enumerate(rows[:10]):
    if "Source" in r:
        hi = i
        break
if hi is None:
    print("no header found; first rows:", rows[:3])
    sys.exit(0)
h = rows[hi]
si = h.index("Source")
cand = [c for c in ("# Samples", "Samples", "Sampling Data (All)", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)") if c in h]
if not cand:
    print("header:", h)
    sys.exit(0)
ss = h.index(cand[0])
body = rows[hi + 1:]
tot = 0
out = []
for j, r in enumerate(body):
    if len(r) <= ss:
        continue
    try:
        n = int(float(r[ss] or 0))
    except ValueError:
        continue
    tot += n
    out.append((n, j))
out.sort(reverse=True)
print("total samples", tot, "column", cand[0])
for n, j in out[:top]:
    for k in range(max(0, j - ctx), j):
        print(f"                 | {body[k][si].strip()[:140]}")
    print(f"{n:7d} {100.0 * n / max(tot, 1):5.1f}%  {body[j][si].strip()[:140]}")
