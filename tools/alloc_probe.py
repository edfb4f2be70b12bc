#!/usr/bin/env python3
"""Developer probe: seconds spent in CUDA start-up, context creation, table upload and photon-buffer allocation."""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marx_b200

t0 = time.time()
lib = marx_b200.load_library()
print("dlopen libmarxb200.so: %.3f s" % (time.time() - t0))
cudart = C.CDLL("libcudart.so.12")
t0 = time.time(); cudart.cudaFree(None); print("cudaFree(0) (CUDA runtime + primary context): %.3f s" % (time.time() - t0))
ctx = C.c_void_p()
t0 = time.time(); assert lib.marxb200_create(C.byref(ctx), 0, 1) == 0; print("marxb200_create: %.3f s" % (time.time() - t0))
t0 = time.time(); assert lib.marxb200_load_calpack(ctx, marx_b200.caldata_path("c2_hetg_acis_s").encode()) == 0
print("marxb200_load_calpack (tables + kernel attribute queries): %.3f s" % (time.time() - t0))
for k in (20, 24):
    t0 = time.time()
    assert lib.marxb200_alloc_photons(ctx, 1 << k) == 0
    print("marxb200_alloc_photons(2^%d = %.2f GB): %.3f s" % (k, (2 * 126 + 36) * (1 << k) / 1e9, time.time() - t0))
t0 = time.time(); assert lib.marxb200_trace(ctx, 0, 1 << 20) == 0; cudart.cudaDeviceSynchronize(); print("first trace of 2^20 rays (module load): %.3f s" % (time.time() - t0))
t0 = time.time(); assert lib.marxb200_trace(ctx, 1 << 20, 1 << 20) == 0; cudart.cudaDeviceSynchronize(); print("second trace: %.4f s" % (time.time() - t0))
lib.marxb200_destroy(ctx)
