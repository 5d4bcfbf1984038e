#!/usr/bin/env python
"""Condense `ncu --set full` captures (gpurun_out/*.ncu-rep) into profiles/r02_ncu_summary.json + a readable table.

usage: python profiles/summarise_ncu.py NAME=path.ncu-rep [NAME=path.ncu-rep ...]
Per kernel: launches captured, average duration, DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum),
FP64 pipe utilisation, issue-slot utilisation, achieved occupancy, registers, shared memory, top stall reasons.
bench.py reads `dram_bytes_per_launch` from the JSON for the `traffic` fields of its roofline objects."""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def col(hdr, units, row, name):
    i = hdr.index(name)
    v = row[i]
    try:
        return float(v) * UNIT.get(units[i], 1.0)
    except ValueError:
        return None


def main():
    summary = {}
    lines = []
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        hdr, units, allrows = load(path)
        kn = hdr.index("Kernel Name")
        if name == "*":          # one entry per kernel found in the report, keyed by its short name
            groups = {}
            for r in allrows:
                groups.setdefault(r[kn].replace("void ", "").split("(")[0].split("<")[0], []).append(r)
        else:
            groups = {name: allrows}
        for name, rows in groups.items():
          n = len(rows)
          avg = lambda m: sum(col(hdr, units, r, m) or 0.0 for r in rows) / n
          stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: sum(float(r[i] or 0) for r in rows) / n
                    for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")}
          top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
          e = {
              "kernel": rows[0][hdr.index("Kernel Name")], "launches_captured": n, "report": os.path.basename(path),
              "duration_us": avg("gpu__time_duration.sum"),
              "dram_bytes_per_launch": avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
              "dram_read_bytes": avg("dram__bytes_read.sum"), "dram_write_bytes": avg("dram__bytes_write.sum"),
              "fp64_pipe_pct_of_peak_active": avg("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
              "fp64_pipe_pct_of_peak_elapsed": avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
              "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
              "registers_per_thread": avg("launch__registers_per_thread"),
              "shared_mem_per_block_bytes": avg("launch__shared_mem_per_block_allocated"),
              "warp_instructions": avg("smsp__inst_executed.sum"),
              "top_stalls_per_issue": {k: round(v, 3) for k, v in top},
          }
          summary[name] = e
          lines.append(f"{name:16s} {e['duration_us']:9.1f} us  dram {e['dram_bytes_per_launch'] / 1e6:8.3f} MB/launch  fp64 pipe "
                       f"{e['fp64_pipe_pct_of_peak_active']:5.1f}% (active) {e['fp64_pipe_pct_of_peak_elapsed']:5.1f}% (elapsed)  issue "
                       f"{e['issue_active_pct']:5.1f}%  warps {e['warps_active_pct']:5.1f}%  regs {e['registers_per_thread']:.0f}  stalls {e['top_stalls_per_issue']}")
    here = os.path.dirname(os.path.abspath(__file__))
    old = {}
    p = os.path.join(here, "r02_ncu_summary.json")
    if os.path.exists(p):
        old = json.load(open(p))
    old.update(summary)
    json.dump(old, open(p, "w"), indent=1)
    open(os.path.join(here, "r02_ncu_summary.txt"), "a").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
