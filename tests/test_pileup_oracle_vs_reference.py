"""Pins oracle/pileup_oracle.c (the plain-C restatement of marxpileup's frame loop, marx/src/marxpileup.c:573-922,1121-1213) to the
UNMODIFIED reference: bit for bit -- every output column, every row -- against the committed output of the stock program run with
counter-based draws (tests/golden/pileup_*.npz, made by oracle/_ref/marxpileup_replay), on fresh simulations where oracle/_ref
exists, and statistically against the stock program with its own RNG.  (The CUDA implementation, marxb200_pileup_run, is compared with this oracle and with
the same fixtures in tests/test_gpu_zz_pileup.py.)"""
import numpy as np
import pytest

from tests import pileup_lib as P


def _same(got, ref, name):
    assert len(got["t"]) == len(ref["t"]) > 0, (name, len(got["t"]), len(ref["t"]))
    for k in ref:
        assert np.array_equal(got[k], ref[k]), (name, k, int((got[k] != ref[k]).sum()))


@pytest.mark.parametrize("name", list(P.CASES))
def test_oracle_reproduces_the_committed_stock_output(name):
    cols, ref, seed = P.load_golden(name)
    got = P.oracle_pileup(cols, P.CASES[name][1], P.CASES[name][2], seed)
    _same(got, ref, name)
    # what the model is: output frames are non-decreasing, every output event carries >= 1 photon, photons are not created
    assert (np.diff(ref["frame"]) >= 0).all() and (ref["nphotons"] >= 1).all()
    assert int(ref["nphotons"].sum()) <= len(cols["t"])
    alpha, ft = P.pileup_params(P.CASES[name][1])
    assert np.array_equal(ref["t"], (ref["frame"] * ft).astype(np.float32))
    # another draw seed changes which piled islands survive grade migration (alpha < 1), never the single-photon events
    other = P.oracle_pileup(cols, P.CASES[name][1], P.CASES[name][2], seed + 1)
    assert (other["nphotons"] == 1).sum() == (ref["nphotons"] == 1).sum()


@pytest.mark.skipif(not P.HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
@pytest.mark.parametrize("name", list(P.CASES))
def test_oracle_reproduces_a_fresh_stock_run(tmp_path, name):
    cols, ref, out = P.stock_case(name, tmp_path, n_rays=120000, seed=17, draw_seed=33)
    got = P.oracle_pileup(cols, P.CASES[name][1], P.CASES[name][2], 33)
    _same(got, ref, name)
    # the stock program with its own generator: same deterministic part, grade migration within binomial noise
    own = P.run_stock_pileup(out, P.CASES[name][1], seed=None)
    assert (own["nphotons"] == 1).sum() == (ref["nphotons"] == 1).sum()
    a, b = int((own["nphotons"] >= 2).sum()), int((ref["nphotons"] >= 2).sum())
    assert abs(a - b) <= 5.0 * np.sqrt(a + b + 1.0), (a, b)
