#!/usr/bin/env python3
"""Generates tests/golden/pileup_*.npz from the UNMODIFIED reference built in oracle/_ref (run in the build container,
`python tests/golden/make_pileup_golden.py`): the stock CPU marx simulates a bright source, oracle/_ref/marxpileup_replay (the
unmodified marxpileup.c with the per-frame Philox stream of oracle/ref/pileup_replay.c) applies the pile-up model.  Stored per case:
the input columns marxpileup read (in.*), the columns it wrote (ref.*), the draw seed."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import pileup_lib as P          # noqa: E402

DRAW_SEED = 9

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(P.CASES)):
        with tempfile.TemporaryDirectory() as d:
            cols, ref, _ = P.stock_case(name, d, n_rays=50000, seed=5, draw_seed=DRAW_SEED)
        blob = {"draw_seed": np.uint64(DRAW_SEED)}
        blob.update({"in." + k: v for k, v in cols.items()})
        blob.update({"ref." + k: v for k, v in ref.items()})
        path = os.path.join(P.GOLDEN, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, len(cols["t"]), "->", len(ref["t"]), "rows ->", path, os.path.getsize(path) // 1024, "KiB")
