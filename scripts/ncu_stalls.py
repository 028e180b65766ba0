#!/usr/bin/env python
"""Print time, issue utilisation, DRAM bytes and the warp-stall breakdown of the first kernel in an .ncu-rep.
usage: python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep [row]"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
v = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
get = lambda k: (v[h.index(k)], u[h.index(k)]) if k in h else ("-", "")
for k in ("Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__throughput.avg.pct_of_peak_sustained_active", "gcc__cache_requests_type_constant.sum.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
    print(f"{k:75s} {get(k)[0][:70]} {get(k)[1]}")
st = []
for i, k in enumerate(h):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        st.append((float(v[i]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
print("stalls (warps per issue):", ", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True) if x >= 0.03))
