#!/usr/bin/env python
"""profiles/conv_traffic.json from an ncu_summary CSV of the 17 conv launches of one forward (variant B, batch 64).
usage: python scripts/make_conv_traffic.py profiles/r1j_conv_full_variantB.csv"""
import csv
import json
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
rows = [r for r in rows if "conv3x3_" in r["kernel"]]
tot = sum(float(r["dram_read_MB"]) + float(r["dram_write_MB"]) for r in rows) * 1e6
out = {"B": {"batch": 64, "launches": len(rows), "dram_bytes_per_step": tot, "dram_bytes_per_launch": tot / len(rows),
             "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum over the "
                       f"{len(rows)} conv3x3_halo / dx / upm / upm2 kernel launches of one forward (scripts/gpu_r2h_conv_evidence.sh -> {sys.argv[1]})"}}
json.dump(out, open("profiles/conv_traffic.json", "w"), indent=1)
print(out)
