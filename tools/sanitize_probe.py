#!/usr/bin/env python3
"""Developer probe for compute-sanitizer: small batches of every configuration through the staged and the fused path,
downloads, the file writer and the packed egress."""
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import marx_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for cfg in ("c2_hetg_acis_s", "c3_letg_hrc_s", "c4_image_acis_i", "c3_hrc_i", "c1_line_acis_s"):
    with marx_b200.MarxB200(cfg, seed=3, max_photons=n) as m:
        m.create_photons(0, n); m.mirror_reflect(); m.grating_diffract(); m.detect()
        a = m.download().copy()
        m.trace(n, n - 7)
        b = m.download().copy()
        d = tempfile.mkdtemp()
        m.write_photons(d, 0x1F01FFF, True, 0.0)
        m.egress_begin_packed(0x1F01FFF, 0.0, n)
        host = np.zeros(n * 120, dtype=np.uint8)
        cols = m.egress_end_packed(host)
        m.set_compaction(False)
        m.create_photons(0, 1000); m.mirror_reflect(); m.grating_diffract(); m.detect()
        c = m.download(all_slots=True)
        m.upload(c); m.detect()
        print(cfg, len(a), len(b), len(cols), len(c), m.stage_counts())
# round 2: consecutive contiguous batches (pre-pass look-ahead, ticketed k01), each followed by the packed egress with the same columns
# (pre-pack: the order restoration writes the file images), then the Level-1-free pile-up paths on the device-resident list
for cfg in ("c2_hetg_acis_s", "c4_beta_acis_i"):
    with marx_b200.MarxB200(cfg, seed=5, max_photons=n) as m:
        host = np.zeros(n * 120, dtype=np.uint8)
        rows = []
        for k in range(4):
            m.trace(k * n, n)
            m.egress_begin_packed(0x1F01FFF, 0.0 if k < 3 else 2.5, n)
            rows.append(len(m.egress_end_packed(host)["energy.dat"]))
        m.trace(9 * n, n // 2)                      # not the predicted batch
        out = m.pileup_events(0.0, 0.5, 3.2, 1, n)
        print(cfg, "egress rows", rows, "pile-up rows", len(out[0]["t"]) if isinstance(out, tuple) else len(out["t"]))
