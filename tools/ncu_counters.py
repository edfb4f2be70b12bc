#!/usr/bin/env python3
"""Per-kernel counters of one traced C2 batch from an `ncu --set full` report, for bench.py's roofline table.

usage: ncu_counters.py REPORT.ncu-rep OUT.json N_GENERATED N_AFTER_A N_AFTER_B1 N_AFTER_B2C1 N_AFTER_MIRROR N_AFTER_GRATING N_DETECTED
(the seven counts are what tools/ncu_probe.py prints: marxb200_get_internal_counts of the profiled batch)

For every kernel class of bench.py's table: executed FP64 flops (2 per DFMA, 1 per DMUL / DADD thread instruction,
smsp__sass_thread_inst_executed_op_d*_pred_on) and DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per INPUT ray of the
class, plus the launch time under the profiler (cold cache, serialised: compare shares, not absolutes)."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
n_gen, n_a, n_b1, n_b2c1, n_mirror, n_grating, n_det = [int(v) for v in sys.argv[3:10]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(r, name):
    return float(r[hdr.index(name)].replace(",", ""))


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


classes = {"K01": ("k01_source_hrma", n_gen), "B1": ("k1_hrma<3>", n_a), "B2C1": ("k1_hrma<4>", n_b1), "C2": ("k1_hrma<5>", n_b2c1),
           "K2": ("k2_", n_mirror), "K3": ("k3_acis", n_grating), "K0": ("k0_time", n_gen), "ORDER": ("order_", n_det)}
res = {}
seen = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].replace("void ", "").replace("mx::", "").replace(" ", "")
    name = name.split("(")[0]
    # the first launch of each distinct kernel only (one traced batch)
    if name in seen:
        continue
    seen[name] = True
    cyc = col(r, "smsp__cycles_elapsed.avg")
    flop = sum(w * col(r, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) * cyc
               for op, w in (("dfma", 2.0), ("dmul", 1.0), ("dadd", 1.0)))
    dram = (to_bytes(col(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")])
            + to_bytes(col(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")]))
    ms = col(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
    for key, (pat, n_in) in classes.items():
        if name.replace(",", ", ").startswith(pat) or name.startswith(pat):
            e = res.setdefault(key, {"kernels": [], "fp64_flop": 0.0, "dram_bytes": 0.0, "ncu_ms": 0.0, "input_rays": n_in})
            e["kernels"].append(name); e["fp64_flop"] += flop; e["dram_bytes"] += dram; e["ncu_ms"] += ms
for key, e in res.items():
    e["fp64_flop_per_input_ray"] = e["fp64_flop"] / max(e["input_rays"], 1)
    e["dram_bytes_per_input_ray"] = e["dram_bytes"] / max(e["input_rays"], 1)
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
for key in sorted(res):
    e = res[key]
    print("%-6s %-40s %8.3f ms  %7.1f flop/ray  %7.1f B/ray" % (key, ",".join(e["kernels"])[:40], e["ncu_ms"], e["fp64_flop_per_input_ray"],
                                                              e["dram_bytes_per_input_ray"]))
