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
print("hostcheck:", "all columns bit-identical" if bad == 0 else "%d mismatches" % bad)
sys.exit(1 if bad else 0)
