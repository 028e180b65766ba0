#!/usr/bin/env python
"""Print the interesting fields of a bench.py JSON line.  usage: python scripts/show_bench.py [file]"""
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_B.json"))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "conv TF", round(d["roofline"]["achieved"], 1), "frac", round(d["roofline"]["frac"], 3))
t = d.get("train_step")
if t:
    print("train ms", round(t["ms_per_step"], 3), "pairs/s", round(t["pairs_per_s"], 1), "launches", t["kernel_launches_per_step"])
for k, v in d["kernels"].items():
    if k != "layers":
        print(" ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
for k, v in d["kernels"].get("layers", {}).items():
    print("   ", k, round(v["ms"], 3), round(v["TFLOPs"] or 0))
for k in ("parity", "gather_check", "cpu_baseline", "clocks"):
    if d.get(k) is not None:
        print(k, d[k])
