"""Statistical parity with the STOCK reference (native RNG, tests/golden/*_stats.npz from oracle/_ref/marx):
detection efficiency (effective area), effective area versus energy, order populations, chip populations, PHA / PI
spectra, chip coordinates and the encircled-energy PSF of the undispersed image must be statistically
indistinguishable (BASELINE.json north_star).  The GPU seeds are fixed, so the test is deterministic."""
import os

import numpy as np
import pytest
from scipy import stats

from tests.stats_bins import summarize

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P_MIN = 1e-4          # per-histogram p-value floor (10 histograms x 4 configs: family-wise ~4e-3)


def two_sample_chi2(a, b):
    """chi^2 test that two histograms with different totals come from the same distribution"""
    a, b = a.astype(np.float64), b.astype(np.float64)
    # merge sparse bins so that every compared bin holds >= 25 counts
    keep = (a + b) >= 25
    if (~keep).any():
        a = np.append(a[keep], a[~keep].sum()); b = np.append(b[keep], b[~keep].sum())
        if a[-1] + b[-1] < 25:
            a, b = a[:-1], b[:-1]
    na, nb = a.sum(), b.sum()
    chi2 = (((np.sqrt(nb / na) * a - np.sqrt(na / nb) * b) ** 2) / (a + b)).sum()
    dof = len(a) - 1
    return chi2, dof, stats.chi2.sf(chi2, dof)


@pytest.mark.parametrize("config", ["c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s", "c4_beta_acis_i"])
def test_distributions_match_stock_marx(config):
    import marx_b200
    ref = np.load(os.path.join(GOLDEN, config + "_stats.npz"))
    n = 1 << 24
    acc, n_gen = None, 0
    with marx_b200.MarxB200(config, seed=20240917, max_photons=n) as m:
        for batch in range(2):
            m.trace(batch * n, n)
            ev = m.download_columns(("energy", "pha", "ccd", "chipx", "chipy", "ypos", "zpos", "shell", "order", "pi"))
            s = summarize(ev)
            acc = s if acc is None else {k: acc[k] + s[k] for k in s}
            n_gen += n
    # detection efficiency = effective area / geometric area (marx.c:597): binomial two-sample z-test
    pa, pb = acc["n_detected"] / n_gen, ref["n_detected"] / ref["n_generated"]
    pool = (acc["n_detected"] + ref["n_detected"]) / (n_gen + ref["n_generated"])
    z = (pa - pb) / np.sqrt(pool * (1 - pool) * (1.0 / n_gen + 1.0 / ref["n_generated"]))
    print(config, "efficiency gpu %.5f ref %.5f z=%.2f" % (pa, pb, z))
    assert abs(z) < 4.0
    report = {}
    for key in ("h_energy", "h_order", "h_ccd", "h_shell", "h_pha", "h_pi", "h_chipx", "h_chipy", "h_psf_r"):
        if key == "h_energy" and config.startswith("c1"):
            continue                                     # monoenergetic
        if key not in ref.files or (acc[key].sum() == 0 and ref[key].sum() == 0):
            continue                                     # column absent for this detector (PI for the HRC)
        chi2, dof, p = two_sample_chi2(acc[key], ref[key])
        report[key] = (round(float(chi2), 1), dof, float(p))
    print(config, report)
    bad = {k: v for k, v in report.items() if v[2] < P_MIN}
    assert not bad, bad
    # encircled energy: KS distance between the two cumulative radial profiles
    ca, cb = np.cumsum(acc["h_psf_r"]) / acc["h_psf_r"].sum(), np.cumsum(ref["h_psf_r"]) / ref["h_psf_r"].sum()
    d = np.abs(ca - cb).max()
    na, nb = acc["h_psf_r"].sum(), ref["h_psf_r"].sum()
    assert d < 1.95 * np.sqrt((na + nb) / (na * nb)), d      # KS critical value at alpha = 0.001
