#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per profiled launch.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.csv"""
import csv
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
        ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem")]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), name, units[hdr.index(k)]) for k, name in WANT if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([name for _, name, _ in cols])
        for r in rows[2:]:
            vals = []
            for i, name, unit in cols:
                v = r[i]
                if name == "kernel":
                    v = v.replace("adn::", "").split("(")[0][:60]
                elif name == "time_us":
                    v = f"{float(v) / (1000.0 if unit == 'ns' else 1.0):.2f}"
                elif name.endswith("_MB"):
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
                    v = f"{float(v) * scale:.2f}"
                elif name.endswith("_pct"):
                    v = f"{float(v):.1f}"
                vals.append(v)
            w.writerow(vals)
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
