#!/usr/bin/env python3
"""Per-kernel summary of the shipped library's device code (cuobjdump): architecture, registers / stack / shared memory, SASS size,
opcode classes (FP64 arithmetic, constant materialisation, TMA bulk copies, tensor-core ops).
usage: python tools/sass_summary.py [LIB] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "marx_b200", "libmarxb200.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage, arch = {}, set()
cur = None
for ln in res.splitlines():
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and cur:
        usage[cur] = tuple(int(v) for v in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = collections.defaultdict(collections.Counter)
cur = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        ops[cur][m.group(1).rstrip(";")] += 1
demangle = subprocess.run(["c++filt"] + list(ops), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(ops, demangle))
print("library: %s   device code: %s   kernels: %d" % (os.path.basename(lib), ", ".join(sorted(arch)), len(ops)))
print("%-46s %5s %6s %7s %7s %6s %6s %6s %6s %7s %7s" % ("kernel", "regs", "stack", "smem", "SASS", "FP64", "UMOV", "LDCU", "MUFU", "UBLKCP", "tensor"))
tot = collections.Counter()
for k in sorted(ops, key=lambda k: names[k]):
    c = ops[k]
    n = sum(v for o, v in c.items() if o != "NOP")
    fp64 = sum(v for o, v in c.items() if o.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"))
    umov = sum(v for o, v in c.items() if o.split(".")[0] == "UMOV")
    ldcu = sum(v for o, v in c.items() if o.split(".")[0] in ("LDCU", "LDC"))
    mufu = sum(v for o, v in c.items() if o.split(".")[0] == "MUFU")
    tma = sum(v for o, v in c.items() if o.startswith("UBLKCP") or o.startswith("UTMA"))
    tens = sum(v for o, v in c.items() if o.split(".")[0] in ("HMMA", "IMMA", "DMMA", "UTCMMA", "UTCHMMA", "QGMMA", "HGMMA"))
    r = usage.get(k, (0, 0, 0, 0))
    short = re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("mx::", "")
    print("%-46s %5d %6d %7d %7d %6d %6d %6d %6d %7d %7d" % (short[:46], r[0], r[1], r[2], n, fp64, umov, ldcu, mufu, tma, tens))
    tot.update({"n": n, "fp64": fp64, "tma": tma, "tens": tens})
print("total SASS instructions %d, FP64 arithmetic %d (%.1f %%), TMA bulk copies (UBLKCP) %d, tensor-core instructions %d"
      % (tot["n"], tot["fp64"], 100.0 * tot["fp64"] / max(tot["n"], 1), tot["tma"], tot["tens"]))
print("(stack = bytes of per-thread local memory reserved for out-of-line calls: the slow paths of FP64 division / square root and the far-range "
      "sin / cos / log fall-backs of mx_math.cuh; smem = static shared memory, the stage kernels add dynamic shared memory at launch)")
