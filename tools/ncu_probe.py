#!/usr/bin/env python3
"""Developer probe for ncu: a few fused marxb200_trace batches (the bench's step) at a reduced batch size."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marx_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 23
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2_hetg_acis_s"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with marx_b200.MarxB200(cfg, seed=1, max_photons=n) as m:
    for r in range(reps):
        m.trace(r * n, n)
    print(m.stage_counts(), m.internal_counts())
