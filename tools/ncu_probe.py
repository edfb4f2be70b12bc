#!/usr/bin/env python3
"""Developer probe for ncu: a few fused marxb200_trace batches (the bench's step), each followed by the Level-1 transform."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import marx_b200
from marx_b200.level1 import Level1Desc

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 23
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2_hetg_acis_s"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "level1_acis_s_hetg_edser.npz"))
desc = {k[5:]: z[k] for k in z.files if k.startswith("desc.")}
with marx_b200.MarxB200(cfg, seed=1, max_photons=n) as m:
    if cfg.endswith("acis_s"):
        m.set_level1(Level1Desc.from_dict(desc))
    for r in range(reps):
        m.trace(r * n, n)
        if cfg.endswith("acis_s"):
            m.level1_transform(0.0)
    print(m.stage_counts(), m.internal_counts())
