#!/usr/bin/env python3
"""Developer probe: fused marxb200_trace batches (the bench's step) with the ABI's per-kernel event timing."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marx_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2_hetg_acis_s"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
stream = torch.cuda.Stream()
with torch.cuda.stream(stream), marx_b200.MarxB200(cfg, seed=1, max_photons=n, stream=stream.cuda_stream) as m:
    for r in range(3):
        m.trace(r * n, n)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r in range(reps):
        m.trace((3 + r) * n, n)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    m.set_profiling(True)
    for r in range(reps):
        m.trace((3 + reps + r) * n, n)
    k = m.kernel_ms()
    print("lib=%s %s n=%d: %.3f ms/batch -> %.3e rays/s | " % (os.path.basename(marx_b200.lib_path()), cfg, n, ms, n / ms * 1e3)
          + "  ".join("%s %.3f" % (a, b[0] / reps) for a, b in k.items() if b[1]), m.stage_counts())
