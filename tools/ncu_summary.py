#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read on the CPU box) into the text committed under profiles/ and refresh
profiles/ncu_traffic.json (the per-launch DRAM traffic bench.py quotes in its `roofline.traffic`).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01c_ncu_full_summary.txt --note "state ..." [--traffic]
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--note", default="")
    ap.add_argument("--traffic", action="store_true", help="rewrite profiles/ncu_traffic.json from this report")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in body:
        name = r[idx["Kernel Name"]]
        seen.setdefault(name, r)  # first captured launch of each kernel (steady state: the capture skips the warm-up)
    lines = ["# ncu --set full --clock-control none, headline shape B=64 F=5000 S=256 ts=4 (bench.py --steps 1 --warmup 3), 1x B200",
             "# " + a.note, "# source report: %s (scratch); one launch of each kernel of a steady-state step" % a.report, ""]
    traffic = {}
    for name, r in seen.items():
        lines.append("## " + name)
        for k in WANT:
            if k in idx:
                lines.append("%-86s %s %s" % (k, r[idx[k]], units[idx[k]]))
        stalls = []
        for h in hdr:
            m = STALL.match(h)
            if m and r[idx[h]]:
                stalls.append((float(r[idx[h]].replace(",", "")), h))
        for v, h in sorted(stalls, reverse=True)[:8]:
            lines.append("%-86s %f inst" % (h, v))
        lines.append("")
        short = re.search(r"(k_\w+)", name)
        if short:
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[k]].replace(",", "")) * UNIT_SCALE.get(units[idx[k]], 1)
            traffic.setdefault(short.group(1), int(tot))
    with open(a.out, "w") as f:
        f.write("\n".join(lines))
    print("\n".join(lines))
    if a.traffic:
        p = os.path.join(os.path.dirname(os.path.abspath(a.out)), "ncu_traffic.json")
        json.dump({"source": "%s (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch, headline "
                             "shape B=64 F=5000 S=256 ts=4)" % a.out, "bytes_per_launch": traffic}, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
