#!/usr/bin/env python3
"""Developer helper: top source lines of a kernel by stall samples / instructions from an .ncu-rep
(needs -lineinfo + --import-source on).  usage: ncu_lines.py REPORT KERNEL_REGEX [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
# kern: "name" (regex on the base name) or "name:N" (N-th launch matching the regex)
name, _, nth = kern.partition(":")
sel = ["--kernel-id", "::regex:%s:%s" % (name, nth)] if nth else ["--kernel-name", "regex:" + name]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, agg = None, None, {}
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    try:
        samples = float(r[hdr.index("# Samples")]); inst = float(r[hdr.index("Instructions Executed")])
        thr = float(r[hdr.index("Thread Instructions Executed")])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0.0, 0.0, 0.0, r[1].strip()[:90]])
    a[0] += samples; a[1] += inst; a[2] += thr
    tot_s += samples; tot_i += inst
print("total samples %.0f, warp instructions %.3e" % (tot_s, tot_i))
print("%-18s %5s %7s %7s %6s  %s" % ("file", "line", "samp%", "inst%", "lanes", "source"))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %5d %7.2f %7.2f %6.1f  %s" % (f, ln, 100 * a[0] / max(tot_s, 1), 100 * a[1] / max(tot_i, 1), a[2] / max(a[1], 1), a[3]))
