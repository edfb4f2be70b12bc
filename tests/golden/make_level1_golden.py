#!/usr/bin/env python3
"""Generates tests/golden/level1_*.npz from the UNMODIFIED reference built in oracle/_ref (run in the build container,
`python tests/golden/make_level1_golden.py`): the stock CPU marx produces an output directory, oracle/_ref/level1_dump
prints the descriptor the stock marx2fits initialisation derives for it, and oracle/_ref/marx2fits_replay (marx2fits.c +
the per-row Philox stream of oracle/ref/level1_rng.c) writes the EVENTS table.  Stored per case: the descriptor, the input
columns marx2fits read (in.*), its EVENTS columns (ref.*), the draw seed."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import level1_lib as L          # noqa: E402

SEED = 11


def make(name, n_rays=100000):
    args, pixadj, ndraw = L.LEVEL1_CASES[name]
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "out")
        L.run_stock_marx(out, args, n_rays=n_rays, seed=7)
        desc = L.dump_descriptor(out, pixadj)
        fits = L.run_stock_marx2fits(out, os.path.join(d, "evt.fits"), pixadj, ndraw, SEED)
        cols = L.read_inputs(out)
    blob = {"seed": np.uint64(SEED), "pixadj": np.array(pixadj), "ndraw": np.int32(ndraw)}
    blob.update({"desc." + k: np.asarray(v) for k, v in desc.items()})
    blob.update({"in." + k: v for k, v in cols.items()})
    blob.update({"ref." + k: v for k, v in fits.items() if k in L.FITS_TO_L1})
    path = os.path.join(L.GOLDEN, name + ".npz")
    np.savez_compressed(path, **blob)
    print(name, len(cols["time"]), "rows ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for case in (sys.argv[1:] or list(L.LEVEL1_CASES)):
        make(case)
