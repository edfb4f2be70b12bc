#!/usr/bin/env python3
"""Statistical reference: runs the STOCK reference driver (oracle/_ref/marx, native JDMrandom stream, energy-sorted
batches, its own output writer) for several seeds and stores summary histograms as a small fixture.  The GPU
path uses a different random stream, so it can only agree with these statistically (chi^2 / KS tests in
tests/test_gpu_statistics.py; BASELINE.json north_star "statistical parity").

    python tests/golden/make_stats.py          (build container only; needs oracle/_ref)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
from tests.stats_bins import summarize  # noqa: E402

from tests.golden.make_golden import COMMON as GOLDEN_COMMON, CONFIGS as GOLDEN_CONFIGS  # noqa: E402

# the same parameter sets the calibration packs / replay fixtures were made with (tests/golden/make_golden.py)
COMMON = GOLDEN_COMMON + ["dNumRays=1000000"]
CONFIGS = {k: GOLDEN_CONFIGS[k]["args"] for k in ("c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s", "c4_beta_acis_i")}
TYPES = {"E": ">f4", "I": ">i2", "J": ">i4", "A": "i1", "D": ">f8"}


def read_dat(path):
    """marx column file (marxio.c:72-77,151-205): 32-byte header (magic[4], type, name[15], nrows i32, ncols i32,
    reserved[4]) followed by big-endian data"""
    b = open(path, "rb").read()
    t = chr(b[4])
    return np.frombuffer(b, TYPES[t], offset=32).astype({"E": "f8", "I": "i4", "J": "i8", "A": "i4", "D": "f8"}[t])


def run_marx(args, seed, nrays, outdir):
    par = os.path.join(outdir, "marx.par")
    subprocess.check_call(["cp", os.path.join(REF, "par", "marx.par"), par])
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    out = os.path.join(outdir, "out")
    subprocess.run([os.path.join(REF, "marx"), "@@" + par, "NumRays=%d" % nrays, "RandomSeed=%d" % seed, "OutputDir=" + out]
                   + COMMON + args, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    cols = {}
    for name, fn in [("energy", "energy.dat"), ("pha", "pha.dat"), ("ccd", "detector.dat"), ("chipx", "xpixel.dat"),
                     ("chipy", "ypixel.dat"), ("ypos", "ypos.dat"), ("zpos", "zpos.dat"), ("shell", "mirror.dat"),
                     ("order", "order.dat"), ("pi", "b_energy.dat")]:
        p = os.path.join(out, fn)
        cols[name] = read_dat(p) if os.path.exists(p) else None
    return cols


def main():
    nrays, seeds = 2000000, [101, 102, 103, 104, 105, 106, 107, 108]
    only = sys.argv[1:]
    for name, args in CONFIGS.items():
        if only and name not in only:
            continue
        acc = None
        n_gen = 0
        for seed in seeds:
            with tempfile.TemporaryDirectory() as d:
                cols = run_marx(args, seed, nrays, d)
            s = summarize(cols)
            acc = s if acc is None else {k: acc[k] + s[k] for k in s}
            n_gen += nrays
        out = os.path.join(ROOT, "tests", "golden", name + "_stats.npz")
        np.savez_compressed(out, n_generated=n_gen, **acc)
        print(name, "generated", n_gen, "detected", int(acc["n_detected"]), "->", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
