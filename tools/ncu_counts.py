#!/usr/bin/env python
"""Per-launch counters of one `ncu --set full` capture -> profiles/ncu_counts.json, the file bench.py quotes (with its
source named) in `roofline_kernels[*].traffic_ncu_capture` and `roofline_issue`: DRAM bytes, warp instructions, the L1
data-stage (LSU wavefront) utilisation, issue utilisation and duration of every kernel of a steady-state step.

    python tools/ncu_counts.py gpurun_out/prof_r2.ncu-rep profiles/ncu_counts.json --source "profiles/r02_ncu_full_summary.txt"
"""
import argparse
import csv
import io
import json
import re
import subprocess

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def get(r, key, scale_unit=False):
        if key not in idx:
            return None
        v = num(r[idx[key]])
        if v is not None and scale_unit:
            v *= UNIT.get(units[idx[key]], 1)
        return v

    kernels = {}
    for r in body:
        m = re.search(r"(k_\w+)", r[idx["Kernel Name"]])
        if not m or m.group(1) in kernels:
            continue  # first captured launch of each kernel (the capture skips the warm-up steps)
        rd, wr = get(r, "dram__bytes_read.sum", True), get(r, "dram__bytes_write.sum", True)
        kernels[m.group(1)] = {
            "dram_bytes": None if rd is None or wr is None else int(rd + wr),
            "warp_instructions": int(get(r, "smsp__inst_executed.sum") or 0),
            "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1_lsu_wavefront_pct": get(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "duration_us_under_ncu": get(r, "gpu__time_duration.sum"),
            "registers": get(r, "launch__registers_per_thread"),
        }
    with open(a.out, "w") as f:
        json.dump({"source": a.source or a.report, "how": "ncu --set full --clock-control none, one launch per kernel of a "
                   "steady-state step of bench.py (headline shape)", "kernels": kernels}, f, indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
