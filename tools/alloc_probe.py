#!/usr/bin/env python3
"""Developer probe: seconds spent creating a context (CUDA start-up + table upload) and allocating the photon buffers of a batch size."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import marx_b200

lib = marx_b200.load_library()
t0 = time.time()
m = marx_b200.MarxB200("c2_hetg_acis_s", seed=1, max_photons=1 << 16)
print("context + tables + 2^16 buffers: %.3f s" % (time.time() - t0))
for k in (20, 22, 23, 24, 25):
    t0 = time.time()
    assert lib.marxb200_alloc_photons(m._ctx, 1 << k) == 0
    print("marxb200_alloc_photons(2^%d = %.2f GB): %.3f s" % (k, (2 * 126 + 36) * (1 << k) / 1e9, time.time() - t0))
m.close()
