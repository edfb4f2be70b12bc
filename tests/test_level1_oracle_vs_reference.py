"""Pins the Level-1 oracle (oracle/level1_oracle.c) to the reference: bit for bit against the committed EVENTS tables of the
stock marx2fits (tests/golden/level1_*.npz) and, where oracle/_ref is built, against fresh runs of the stock programs with
other seeds; plus the batch-carry property (a file transformed in pieces equals the file transformed at once)."""
import os

import numpy as np
import pytest

from tests import level1_lib as L

CASES = sorted(L.LEVEL1_CASES)


@pytest.mark.parametrize("case", CASES)
def test_level1_oracle_reproduces_the_stock_marx2fits_table(case):
    desc, cols, ref, seed = L.load_golden(case)
    l1 = L.Level1Oracle(desc, seed).transform(cols)
    report = L.compare_with_fits(l1, ref, f32_ulps=0)
    assert len(report) >= 12 and len(cols["time"]) > 2000


@pytest.mark.parametrize("case", CASES)
def test_level1_oracle_batches_carry_the_exposure_state(case):
    desc, cols, _, seed = L.load_golden(case)
    whole = L.Level1Oracle(desc, seed).transform(cols)
    o = L.Level1Oracle(desc, seed)
    n = len(cols["time"])
    cuts = [0, 1, n // 3, n // 3 + 1, (2 * n) // 3, n]
    parts = [o.transform({k: v[a:b] for k, v in cols.items()}) for a, b in zip(cuts[:-1], cuts[1:])]
    for k in whole:
        assert np.array_equal(whole[k], np.concatenate([p[k] for p in parts])), k


def test_level1_frames_share_the_aspect_of_their_first_event():
    """read_dither_value (marx2fits.c:3567-3580): perturbing the aspect of an event that is not the first of its exposure frame
    changes nothing (ACIS, not --pixadj=exact); perturbing a frame's first event moves the whole frame."""
    desc, cols, _, seed = L.load_golden("level1_acis_s_hetg_edser")
    base = L.Level1Oracle(desc, seed).transform(cols)
    expno = base["expno"]
    first = np.flatnonzero(np.r_[True, expno[1:] != expno[:-1]])
    later = np.flatnonzero(np.r_[False, expno[1:] == expno[:-1]])
    assert len(later) > 100
    c2 = {k: v.copy() for k, v in cols.items()}
    c2["sky_ra"][later] += 1e-4
    moved = L.Level1Oracle(desc, seed).transform(c2)
    assert np.array_equal(moved["x"], base["x"]) and np.array_equal(moved["y"], base["y"])
    c3 = {k: v.copy() for k, v in cols.items()}
    c3["sky_ra"][first] += 1e-4
    moved = L.Level1Oracle(desc, seed).transform(c3)
    assert (moved["x"] != base["x"]).mean() > 0.99


@pytest.mark.skipif(not L.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case,seed,rays", [("level1_acis_s_hetg_edser", 3, 300000), ("level1_acis_i_beta_randomize", 4, 100000),
                                            ("level1_hrc_s_letg", 5, 200000), ("level1_acis_s_hetg_exact", 6, 100000),
                                            ("level1_acis_s_nodither_none", 8, 100000)])
def test_level1_oracle_against_a_fresh_reference_run(case, seed, rays, tmp_path):
    args, pixadj, ndraw = L.LEVEL1_CASES[case]
    out = tmp_path / "out"
    L.run_stock_marx(out, args, n_rays=rays, seed=seed)
    desc = L.dump_descriptor(out, pixadj)
    fits = L.run_stock_marx2fits(out, tmp_path / "evt.fits", pixadj, ndraw, seed=100 + seed)
    l1 = L.Level1Oracle(desc, 100 + seed).transform(L.read_inputs(out))
    L.compare_with_fits(l1, fits, f32_ulps=0)
