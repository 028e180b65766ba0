#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/launch_agg.py launches.csv [n_steps_in_file] [top]   (the last 1/n_steps of the launches = one step)"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
data = [(r[ki], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0 if r[ui] == "us" else 1e3)) for r in rows[hi + 1:] if len(r) > vi]
n = len(data) // steps
agg = collections.defaultdict(lambda: [0, 0.0])
for k, t in data[-n:]:
    k = re.sub(r"\(.*", "", k).replace("adn::", "").replace("void ", "")
    agg[k][0] += 1; agg[k][1] += t
tot = sum(v[1] for v in agg.values())
print(f"launches/step {n}, sum of kernel durations {tot:.0f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[:72]:72s} {v[0]:4d} {v[1]:9.1f} us {100 * v[1] / tot:5.1f}%")
