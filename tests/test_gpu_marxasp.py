"""The aspect-solution table on the GPU (marxb200_aspsol_rows = the row loop of marxasp, marx/src/marxasp.c:996-1027) against the
stock marxasp tables (committed windows, tests/golden/aspsol_*.npz; complete fresh tables where oracle/_ref exists) and the plain-C
restatement, which tests/test_aspsol_oracle_vs_reference.py pins bit-exact to the stock program.

Tolerances (FP64; libdevice's sin / cos / acos / asin differ from the host libm by an ulp or two):
  TIME exact; RA, DEC, ROLL within 1e-9 relative (north_star) -- measured ~1e-15; RA additionally within the conditioning bound
  of acos / asin near the pole (|dRA| * cos(DEC) <= 1e-11 deg); quaternion components within 1e-12 absolute.
The FITS row image must hold exactly the column values, big endian, in the table's layout (4D 3E 4D = 76 bytes)."""
import numpy as np
import pytest

import marx_b200
from tests import aspsol_lib as A

pytestmark = pytest.mark.gpu

ROW = np.dtype([("time", ">f8"), ("ra", ">f8"), ("dec", ">f8"), ("roll", ">f8"), ("dy", ">f4"), ("dz", ">f4"), ("dtheta", ">f4"),
                ("q_att", ">f8", (4,))])


def compare(got, ref, name):
    assert np.array_equal(got["time"], ref["time"]), name
    cosd = np.maximum(np.cos(np.radians(ref["dec"])), 1e-6)
    d_ra = np.abs(got["ra"] - ref["ra"])
    d_ra = np.minimum(d_ra, 360.0 - d_ra)
    assert (d_ra * cosd).max() <= 1e-11, (name, "ra", (d_ra * cosd).max())
    assert np.abs(got["dec"] - ref["dec"]).max() <= 1e-9 * 90.0, (name, "dec")
    assert (np.abs(got["roll"] - ref["roll"]) <= 1e-9 * np.maximum(np.abs(ref["roll"]), 1.0)).all(), (name, "roll")
    for j in range(4):
        assert np.abs(got["q%d" % j] - ref["q%d" % j]).max() <= 1e-12, (name, "q%d" % j)
    return {"ra*cos(dec)": float((d_ra * cosd).max()), "dec": float(np.abs(got["dec"] - ref["dec"]).max()),
            "roll": float(np.abs(got["roll"] - ref["roll"]).max()),
            "q": float(max(np.abs(got["q%d" % j] - ref["q%d" % j]).max() for j in range(4)))}


@pytest.mark.parametrize("name", list(A.CASES))
def test_rows_match_the_committed_stock_table(name):
    desc, num, first, ref = A.load_golden(name)
    n = len(ref["time"])
    with marx_b200.MarxB200("c1_acis_s", max_photons=1024) as m:
        got, img, ms = m.aspsol_rows(desc, first, n, fits_rows=True)
    print(name, compare(got, ref, name))
    ora = A.oracle_rows(desc, first, n)
    compare(got, ora, name + " (oracle)")
    rows = img.view(ROW)
    assert len(rows) == n
    for k in ("time", "ra", "dec", "roll"):
        assert np.array_equal(rows[k].astype(np.float64), got[k]), k
    for j in range(4):
        assert np.array_equal(rows["q_att"][:, j].astype(np.float64), got["q%d" % j])
    assert (rows["dy"] == 0).all() and (rows["dz"] == 0).all() and (rows["dtheta"] == 0).all()


@pytest.mark.skipif(not A.HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
def test_complete_fresh_table_and_row_windows(tmp_path):
    """a whole table of the stock marxasp (other seed than the fixtures), computed in one call and in ragged windows"""
    name = "aspsol_roll_pole"
    desc, num, ref = A.stock_case(name, tmp_path, n_rays=12000, seed=21)
    with marx_b200.MarxB200("c1_acis_s", max_photons=1024) as m:
        got, _, _ = m.aspsol_rows(desc, 0, num)
        print(name, num, compare(got, ref, name))
        for first, n in ((0, 1), (255, 2), (num - 300, 300), (1000, 257)):
            part, img, _ = m.aspsol_rows(desc, first, n, fits_rows=True)
            for k in A.COLS:
                assert np.array_equal(part[k], got[k][first:first + n]), (first, n, k)
            assert np.array_equal(img.view(ROW)["ra"].astype(np.float64), part["ra"])
        empty, _, _ = m.aspsol_rows(desc, 5, 0)
        assert len(empty["time"]) == 0


def test_one_million_rows_properties_and_rate():
    """size-independent properties at 2^20 rows (3 days of exposure at 0.256 s): unit quaternions, dither amplitudes, the
    table equals the oracle on a strided sample; prints the device rate next to the one-core rate of the restatement"""
    import time
    desc, _, _, _ = A.load_golden("aspsol_default")
    n = 1 << 20
    with marx_b200.MarxB200("c1_acis_s", max_photons=1024) as m:
        m.aspsol_rows(desc, 0, n)
        got, img, ms = m.aspsol_rows(desc, 0, n, fits_rows=True)
    q = np.stack([got["q%d" % j] for j in range(4)])
    assert np.abs((q * q).sum(axis=0) - 1.0).max() < 1e-15
    assert np.array_equal(got["time"], np.arange(n) * desc[1] + desc[0])
    # the RA / Dec offsets are rolled about the pointing by the nominal roll (marxasp.c:849-851): the Dec excursion mixes both
    amp_deg = np.degrees(np.hypot(desc[2], desc[3]))
    assert 0.5 * amp_deg < np.ptp(got["dec"]) <= 2.0 * amp_deg * (1 + 1e-9)
    assert len(img) == 76 * n
    t0 = time.time()
    ora = A.oracle_rows(desc, 0, 1 << 17)
    cpu = (1 << 17) / (time.time() - t0)
    sample = {k: got[k][:1 << 17] for k in A.COLS}
    compare(sample, ora, "2^17 of 2^20")
    print("aspsol rows: %.3f ms for %d rows on the device = %.3e rows/s; oracle (one core) %.3e rows/s" % (ms, n, n / ms * 1e3, cpu))
