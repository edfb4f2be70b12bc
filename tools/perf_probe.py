#!/usr/bin/env python3
"""Developer probe: per-stage device times (CUDA events on the library's stream) for one batch."""
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marx_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2_hetg_acis_s"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    m = marx_b200.MarxB200(cfg, seed=1, max_photons=n, stream=stream.cuda_stream)
    names = ["create", "mirror", "grating", "detect", "order"]
    calls = [lambda i: m.create_photons(i * n, n), m.mirror_reflect, m.grating_diffract, m.detect, m.restore_order]
    for rep in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(stream)
        for k, c in enumerate(calls):
            if k == 0:
                c(rep)
            else:
                c()
            ev[k + 1].record(stream)
        stream.synchronize()
        ts = [ev[k].elapsed_time(ev[k + 1]) for k in range(5)]
        tot = sum(ts)
        print("rep %d: " % rep + "  ".join("%s %.3f ms" % (a, b) for a, b in zip(names, ts)),
              " total %.3f ms  -> %.3e rays/s" % (tot, n / tot * 1e3), m.stage_counts())
