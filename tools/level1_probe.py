#!/usr/bin/env python3
"""Developer probe: device time of marxb200_level1_transform on the events of one traced batch."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marx_b200
from marx_b200.level1 import Level1Desc, PIXADJ
from tests import level1_lib as L

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2_hetg_acis_s"
fixture = sys.argv[3] if len(sys.argv) > 3 else "level1_acis_s_hetg_edser"
pixadj = sys.argv[4] if len(sys.argv) > 4 else None
reps = 20
desc = L.load_golden(fixture)[0]
if pixadj:
    desc = dict(desc, pix_adjust=PIXADJ[pixadj])
stream = torch.cuda.Stream()
with torch.cuda.stream(stream), marx_b200.MarxB200(cfg, seed=1, max_photons=n, stream=stream.cuda_stream) as m:
    m.set_level1(Level1Desc.from_dict(desc))
    m.trace(0, n)
    events = m.counts()[1]
    for _ in range(3):
        m.level1_transform(0.0)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        m.level1_transform(0.0)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("level1 %s/%s pixadj=%s: %d events, %.4f ms per transform (incl. the host count read) -> %.3e events/s" %
          (cfg, fixture, pixadj or "fixture", events, ms, events / ms * 1e3))
    m.set_profiling(True)
    for _ in range(reps):
        m.level1_transform(0.0)
    k = m.kernel_ms()["level1"]
    print("  kernels only: %.4f ms -> %.3e events/s" % (k[0] / k[1], events / (k[0] / k[1]) * 1e3))
