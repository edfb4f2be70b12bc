"""Developer harness: run tools/hostcheck/pileup_hostcheck (the MX_HD steps of mx_pileup.cuh on the host) on the committed
pile-up fixtures and compare every output column with the stock program's, bit for bit.  Not a test, not a compute path."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.pileup_lib import CASES, load_golden, pileup_params  # noqa: E402

exe = os.path.join(ROOT, "build", "pileup_hostcheck")
os.makedirs(os.path.dirname(exe), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", "-x", "c++",
                       os.path.join(ROOT, "tools/hostcheck/pileup_hostcheck.cpp"), os.path.join(ROOT, "marx_b200/csrc/calpack.cpp"), "-o", exe])
bad = 0
for name, (_, pu_args, pack) in CASES.items():
    cols, ref, seed = load_golden(name)
    alpha, ft = pileup_params(pu_args)
    with tempfile.TemporaryDirectory() as d:
        for k, v in cols.items():
            np.ascontiguousarray(v).tofile(os.path.join(d, "in.%s.bin" % k))
        out = subprocess.run([exe, os.path.join(ROOT, "marx_b200/caldata", pack + ".calpack"), d, str(len(cols["t"])), repr(alpha), repr(ft), str(seed)],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(name, out.stdout.strip())
        if out.returncode != 0:
            bad += 1
            continue
        for k, r in ref.items():
            g = np.fromfile(os.path.join(d, "out.%s.bin" % k), dtype=r.dtype)
            same = (g.shape == r.shape) and (g.tobytes() == r.tobytes())
            if not same:
                bad += 1
                m = min(len(g), len(r))
                diff = np.nonzero(g[:m].view(np.uint8).reshape(m, -1) != r[:m].view(np.uint8).reshape(m, -1))[0]
                print("   MISMATCH %-10s rows %d vs %d, first differing row %s" % (k, len(g), len(r), diff[:1]))
# the synthetic event lists of tests/test_gpu_zz_pileup.py (heavy pile-up, several chips, border events) against the pinned oracle
if "--synthetic" in sys.argv:
    from tests.pileup_lib import oracle_pileup  # noqa: E402
    from tests.test_gpu_zz_pileup import _synthetic  # noqa: E402
    for n, rate, alpha, ft, spot, ccd in [(200000, 40.0, 0.5, 3.241, 6.0, (7,)), (300000, 2.0, 0.9, 3.2, 1.5, (7,)),
                                          (100000, 500.0, 0.2, 0.4, 30.0, (5, 6, 7, 8)), (50000, 3000.0, 1.0, 3.2, 40.0, (7,))]:
        cols = _synthetic(n, rate, 1234 + n, spot=spot, ccd=ccd)
        ref = oracle_pileup(cols, ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"], "c1_acis_s", 77)
        with tempfile.TemporaryDirectory() as d:
            for k, v in cols.items():
                np.ascontiguousarray(v).tofile(os.path.join(d, "in.%s.bin" % k))
            out = subprocess.run([exe, os.path.join(ROOT, "marx_b200/caldata", "c1_acis_s.calpack"), d, str(n), repr(alpha), repr(ft), "77"],
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print("synthetic", n, rate, out.stdout.strip(), "piled rows", int((ref["nphotons"] >= 2).sum()))
            if out.returncode != 0:
                bad += 1
                continue
            for k, r in ref.items():
                g = np.fromfile(os.path.join(d, "out.%s.bin" % k), dtype=r.dtype)
                if g.tobytes() != r.tobytes():
                    bad += 1
                    print("   MISMATCH", k, len(g), len(r))
print("hostcheck:", "all columns bit-identical" if bad == 0 else "%d mismatches" % bad)
sys.exit(1 if bad else 0)
