#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line (runs on the CPU box, no GPU).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_edge_scan [--so neural_renderer_b200/libnr_b200.so] [--top 30]

`ncu --page source --csv` lists SASS instructions with executed counts and stall samples but no source lines; the
line table comes from `nvdisasm -g` on the cubin of the SAME build (the library must have been compiled with
-lineinfo and must be the one that was profiled).  Instructions are matched by order within the kernel.
"""
import argparse
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, kernel_substr, template=None):
    """[(sass text, file, line)] in program order for the first kernel whose mangled name contains kernel_substr."""
    out = []
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for cubin in sorted(glob.glob(os.path.join(td, "*.cubin"))):
            txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
            cur = None
            fl = ("?", 0)
            for ln in txt.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", ln)
                if m:
                    name = m.group(1)
                    cur = name if (kernel_substr in name and (template is None or template in name)) else None
                    if cur and out:
                        return out
                    continue
                if cur is None:
                    continue
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if m:
                    fl = (m.group(1), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    out.append((m.group(2).strip(), fl[0], fl[1]))
            if out:
                return out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--template", default=None, help="extra substring of the mangled name, e.g. ILi1E")
    ap.add_argument("--so", default="neural_renderer_b200/libnr_b200.so")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--launch", type=int, default=0, help="index among the launches of this kernel in the report")
    ap.add_argument("--phases", action="store_true",
                    help="also aggregate by the `//@phase <label>` comment markers of the source files (a marker labels "
                         "every line up to the next marker of the same file)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the CSV holds one block per launch: "Kernel Name",<name> / header / rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if len(row) >= 2 and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = row
            else:
                cur["rows"].append(row)
    blocks = [b for b in blocks if a.kernel in b["name"]]
    if not blocks:
        sys.exit("kernel not in report")
    b = blocks[a.launch]
    idx = {h: i for i, h in enumerate(b["hdr"])}
    lines = sass_lines(a.so, a.kernel, a.template)
    if len(lines) != len(b["rows"]):
        print("warning: %d SASS instructions in the library vs %d in the report (different build?)"
              % (len(lines), len(b["rows"])), file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for k, row in enumerate(b["rows"]):
        ie = int(row[idx["Instructions Executed"]] or 0)
        ti = int(row[idx["Thread Instructions Executed"]] or 0)
        sm = int(row[idx["# Samples"]] or 0)
        f, l = (lines[k][1], lines[k][2]) if k < len(lines) else ("?", 0)
        g = agg[(os.path.basename(f), l)]
        g[0] += ie
        g[1] += ti
        g[2] += sm
        tot_i += ie
        tot_s += sm
    print("kernel: %s\nwarp instructions %d, stall samples %d" % (b["name"], tot_i, tot_s))
    src_cache = {}

    def src(f, l):
        for root in (".", "neural_renderer_b200/csrc"):
            p = os.path.join(root, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < l <= len(src_cache[p]):
                    return src_cache[p][l - 1].strip()[:90]
        return ""

    if a.phases:
        markers = {}  # file -> sorted [(line, label)]

        def phase_of(f, l):
            if f not in markers:
                ms = []
                for root in (".", "neural_renderer_b200/csrc"):
                    pth = os.path.join(root, f)
                    if os.path.exists(pth):
                        for k, text in enumerate(open(pth).read().splitlines(), 1):
                            m = re.search(r"//@phase\s+(.*)", text)
                            if m:
                                ms.append((k, m.group(1).strip()))
                        break
                markers[f] = ms
            label = "(unmarked)"
            for k, lab in markers[f]:
                if k <= l:
                    label = lab
                else:
                    break
            return "%s: %s" % (f, label)

        ph = defaultdict(lambda: [0, 0])
        for (f, l), (ie, ti, sm) in agg.items():
            g = ph[phase_of(f, l)]
            g[0] += ie
            g[1] += sm
        print("%-78s %8s %8s" % ("phase", "inst %", "samp %"))
        for name, (ie, sm) in sorted(ph.items(), key=lambda kv: -kv[1][0]):
            print("%-78s %7.2f%% %7.2f%%" % (name[:78], 100.0 * ie / max(tot_i, 1), 100.0 * sm / max(tot_s, 1)))
        print()
    print("%-22s %8s %8s %6s  %s" % ("file:line", "inst %", "samp %", "lanes", "source"))
    for (f, l), (ie, ti, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print("%-22s %7.2f%% %7.2f%% %6.1f  %s" % ("%s:%d" % (f, l), 100.0 * ie / max(tot_i, 1), 100.0 * sm / max(tot_s, 1),
                                                 ti / max(ie, 1), src(f, l)))


if __name__ == "__main__":
    main()
