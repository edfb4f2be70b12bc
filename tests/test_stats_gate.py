"""CPU checks of the statistical gate used by tests/test_gpu_statistics.py (Holm-Bonferroni step-down, two-sample chi^2)."""
import numpy as np

from tests.test_gpu_statistics import ALPHA, holm_failures, two_sample_chi2


def test_holm_step_down():
    assert holm_failures({"a": 0.5, "b": 0.2, "c": 0.011}, 0.03) == {}                  # smallest threshold 0.01
    bad = holm_failures({"a": 0.5, "b": 0.012, "c": 0.009}, 0.03)                       # c < 0.01 rejected, then b < 0.015 rejected
    assert set(bad) == {"b", "c"}
    assert set(holm_failures({"a": 0.004, "b": 0.9}, ALPHA)) == {"a"}                   # 0.004 < 0.01 / 2


def test_two_sample_chi2_is_calibrated():
    r = np.random.default_rng(5)
    pr = r.dirichlet(np.ones(40))
    ps = [two_sample_chi2(r.multinomial(200000, pr), r.multinomial(70000, pr))[2] for _ in range(400)]
    assert 0.002 < np.mean(np.array(ps) < 0.01) + 0.005 and np.mean(np.array(ps) < 0.01) < 0.04     # ~1 % false alarms
    assert abs(np.mean(ps) - 0.5) < 0.06
    q = pr.copy(); q[:5] *= 1.15; q /= q.sum()
    assert two_sample_chi2(r.multinomial(200000, pr), r.multinomial(70000, q))[2] < 1e-6             # and it has power
