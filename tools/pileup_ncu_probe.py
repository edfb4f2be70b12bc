#!/usr/bin/env python3
"""Developer probe for ncu: one marxb200_pileup_run on the bench's synthetic list (2^22 events, ~32 per frame)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import marx_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 22
r = np.random.default_rng(1)
alpha, ft, rate = 0.5, 3.241, 10.0
cols = {"ccd": np.full(n, 7, np.int8), "t": np.cumsum(r.exponential(1.0 / rate, n)).astype(np.float32),
        "x": (512.0 + r.normal(0.0, 3.0, n)).astype(np.float32), "y": (300.0 + r.normal(0.0, 3.0, n)).astype(np.float32),
        "benergy": r.uniform(0.4, 7.0, n).astype(np.float32)}
for k in ("sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta"):
    cols[k] = r.normal(0.0, 1e-3, n).astype(np.float32)
with marx_b200.MarxB200("c2_hetg_acis_s", seed=1, max_photons=1024) as m:
    for _ in range(2):
        got, ms = m.pileup(cols, alpha, ft, 1)
    print("rows", len(got["t"]), "kernel ms", ms)
