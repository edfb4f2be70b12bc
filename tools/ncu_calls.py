#!/usr/bin/env python3
"""Developer helper: the out-of-line CALLs a kernel actually executes, per source line, from an .ncu-rep captured with
-lineinfo + --import-source on.  usage: ncu_calls.py REPORT KERNEL_REGEX[:N] [min_calls]

Why: sm_100's FP64 division, sqrt, pow, sin ... have inlined fast paths and out-of-line special-case routines (~100 instructions).  A
quotient that is exactly zero, a denormal operand or a huge argument takes the routine; when that happens for every ray the source line
looks innocent in the line table (the callee's instructions are attributed to some other line with debug info), but the CALL instruction
itself carries the execution count.  Round 2 found four such divisions this way."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
min_calls = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
name, _, nth = kern.partition(":")
sel = ["--kernel-id", "::regex:%s:%s" % (name, nth)] if nth else ["--kernel-name", "regex:" + name]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, cur, hdr, found, total = None, None, None, {}, 0
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    i_sass = [k for k, h in enumerate(hdr) if h == "Source"][1]
    i_exec, i_thr = hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    if r[0].strip().isdigit():
        cur = (fname, r[0].strip(), r[1].strip()[:100])
    if not r[i_exec].isdigit():
        continue
    total += int(r[i_exec])
    if "CALL" in r[i_sass] and int(r[i_exec]) >= min_calls:
        found[(cur, r[i_sass].split()[-1])] = (int(r[i_exec]), r[i_thr])
print("%s: %d CALL sites executed >= %d times (warp level); the listing holds each instruction once per inlining context" % (kern, len(found), min_calls))
for (src, target), (n, lanes) in sorted(found.items(), key=lambda kv: -kv[1][0]):
    print("  %8d calls, %4s lanes  %s:%s  %s   -> %s" % (n, lanes, src[0], src[1], src[2], target))
