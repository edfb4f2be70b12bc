#!/usr/bin/env python3
"""Generates tests/golden/aspsol_*.npz from the UNMODIFIED reference built in oracle/_ref (run in the build container,
`python tests/golden/make_aspsol_golden.py`): the stock CPU marx produces an output directory, oracle/_ref/asp_dump prints the
descriptor the stock marxasp initialisation derives for it, oracle/_ref/marxasp writes the ASPSOL table.  Stored per case: the
descriptor, the row count, and a window of 3000 rows of the table (time, ra, dec, roll, q_att) starting at `first_row`."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import aspsol_lib as A          # noqa: E402

WINDOW = 3000

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(A.CASES)):
        with tempfile.TemporaryDirectory() as d:
            desc, num, ref = A.stock_case(name, d)
        assert len(ref["time"]) == num and (ref["dy"] == 0).all() and (ref["dz"] == 0).all() and (ref["dtheta"] == 0).all()
        first = max(0, min(num - WINDOW, 1000))
        blob = {"desc": desc, "num_rows": np.int64(num), "first_row": np.int64(first)}
        blob.update({"ref." + k: ref[k][first:first + WINDOW] for k in A.COLS})
        path = os.path.join(A.GOLDEN, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, num, "rows ->", path, os.path.getsize(path) // 1024, "KiB")
