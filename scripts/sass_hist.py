#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass` dump: executed warp-instructions per opcode, and the
top stall-sample addresses.  usage: python scripts/sass_hist.py dump.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
si, ei, ss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops = collections.Counter()
samp = collections.Counter()
total = 0
lines = []
for r in rows[2:]:
    if len(r) <= ei:
        continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    n = int(r[ei] or 0)
    ops[op] += n
    samp[op] += int(r[ss] or 0)
    total += n
    lines.append((int(r[ss] or 0), n, src))
print("total warp-instructions executed:", total)
for op, n in ops.most_common(top):
    print(f"{op:10s} {n:12d} {100.0 * n / total:5.1f}%   samples {samp[op]}")
print("--- top stall-sample lines")
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(s, n, src)
