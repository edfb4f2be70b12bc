"""Pins oracle/aspsol_oracle.c (the plain-C restatement of marxasp's row loop, marx/src/marxasp.c:814-903,996-1027) to the
UNMODIFIED reference: bit for bit against the committed windows of stock ASPSOL tables (tests/golden/aspsol_*.npz) and, where
oracle/_ref exists, against complete fresh tables of the stock marxasp with other seeds / exposure lengths."""
import numpy as np
import pytest

from tests import aspsol_lib as A


@pytest.mark.parametrize("name", list(A.CASES))
def test_oracle_reproduces_the_committed_stock_table(name):
    desc, num, first, ref = A.load_golden(name)
    n = len(ref["time"])
    got = A.oracle_rows(desc, first, n)
    for k in A.COLS:
        assert np.array_equal(got[k], ref[k]), (name, k, np.abs(got[k] - ref[k]).max())
    # the table is what the descriptor says: uniform sampling from TSTART, unit quaternions
    assert np.array_equal(ref["time"], (np.arange(first, first + n) * desc[1]) + desc[0])
    q = np.stack([ref["q%d" % j] for j in range(4)])
    assert np.abs((q * q).sum(axis=0) - 1.0).max() < 1e-15


@pytest.mark.skipif(not A.HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
@pytest.mark.parametrize("name", list(A.CASES))
def test_oracle_reproduces_a_fresh_stock_table(tmp_path, name):
    desc, num, ref = A.stock_case(name, tmp_path, n_rays=8000, seed=11)
    assert len(ref["time"]) == num > 100
    got = A.oracle_rows(desc, 0, num)
    for k in A.COLS:
        assert np.array_equal(got[k], ref[k]), (name, k)
    assert (ref["dy"] == 0).all() and (ref["dz"] == 0).all() and (ref["dtheta"] == 0).all()
