"""ACIS pile-up on the GPU (marxb200_pileup_run / marxb200_pileup_events; marx/src/marxpileup.c:573-922,1121-1213; SURVEY.md 8f
rank 4) against the stock program's committed output (tests/golden/pileup_*.npz, bit for bit, every column and row) and against the
pinned plain-C oracle (oracle/pileup_oracle.c) on larger and on adversarial event lists.  Both device forms are covered: the fused
single kernel (frames staged in shared memory) and the eight step kernels it falls back to for frames beyond its window."""
import os
import numpy as np
import pytest

import marx_b200
from tests import pileup_lib as P

pytestmark = pytest.mark.gpu
PACK = {name: marx_b200.caldata_path(P.CASES[name][2] + ".calpack") for name in P.CASES}


def _same(got, ref, what):
    assert len(got["t"]) == len(ref["t"]), (what, len(got["t"]), len(ref["t"]))
    for k in ref:
        assert k in got, (what, k)
        assert got[k].dtype == ref[k].dtype and got[k].tobytes() == ref[k].tobytes(), (what, k, int((got[k] != ref[k]).sum()))


@pytest.mark.parametrize("name", list(P.CASES))
def test_device_reproduces_the_committed_stock_output(name):
    cols, ref, seed = P.load_golden(name)
    alpha, ft = P.pileup_params(P.CASES[name][1])
    with marx_b200.MarxB200(PACK[name], seed=1, max_photons=1024) as m:
        before = m.launch_count()
        got, ms = m.pileup(cols, alpha, ft, seed)
        assert m.launch_count() - before == 4 and ms > 0.0           # the fused kernel + the three row-placement kernels
        _same(got, ref, name)
        os.environ["MARXB200_PILEUP_FUSED"] = "0"                     # the eight step kernels
        try:
            before = m.launch_count()
            steps, _ = m.pileup(cols, alpha, ft, seed)
            assert m.launch_count() - before == 8
        finally:
            del os.environ["MARXB200_PILEUP_FUSED"]
        _same(steps, ref, name + " step kernels")
        # frames are independent: any split of the list at a frame boundary gives the same rows
        frame = (cols["t"].astype(np.float64) / ft).astype(np.uint32)
        cut = int(np.nonzero(np.diff(frame))[0][len(np.nonzero(np.diff(frame))[0]) // 2]) + 1
        a, _ = m.pileup({k: v[:cut] for k, v in cols.items()}, alpha, ft, seed)
        b, _ = m.pileup({k: v[cut:] for k, v in cols.items()}, alpha, ft, seed)
        _same({k: np.concatenate([a[k], b[k]]) for k in a}, ref, name + " split")
        # without the dither columns (the stock program cannot read such a directory; the oracle can)
        bare = {k: cols[k] for k in ("ccd", "x", "y", "t", "benergy")}
        got2, _ = m.pileup(bare, alpha, ft, seed)
        _same(got2, {k: ref[k] for k in got2}, name + " bare")


def _synthetic(n, rate, seed, spot=6.0, ccd=(7,), edge=0.02):
    """a bright spot dithered over a few pixels, `rate` events per second, a fraction `edge` of the events on the chip border"""
    r = np.random.default_rng(seed)
    t = np.cumsum(r.exponential(1.0 / rate, n)).astype(np.float32)
    x = (512.0 + r.normal(0.0, spot, n)).astype(np.float32)
    y = (300.0 + r.normal(0.0, spot, n)).astype(np.float32)
    on_edge = r.random(n) < edge
    x[on_edge] = r.choice(np.array([0.3, 0.99, 1023.0, 1023.7], np.float32), int(on_edge.sum()))
    return {"ccd": r.choice(np.array(ccd, np.int8), n), "x": x, "y": y, "t": t, "benergy": r.uniform(0.4, 7.0, n).astype(np.float32),
            **{k: r.normal(0.0, 1e-3, n).astype(np.float32) for k in P.DITHER}}


@pytest.mark.parametrize("n,rate,alpha,ft,spot,ccd", [
    (200000, 40.0, 0.5, 3.241, 6.0, (7,)),            # ~130 events per frame on a few hundred pixels: heavy pile-up, many duplicates
    (300000, 2.0, 0.9, 3.2, 1.5, (7,)),               # ~6 events per frame on a tight spot
    (100000, 500.0, 0.2, 0.4, 30.0, (5, 6, 7, 8)),    # four chips share the frames
    (50000, 3000.0, 1.0, 3.2, 40.0, (7,)),            # ~10^4 events per frame, alpha = 1: every island survives
])
def test_device_equals_the_oracle_on_synthetic_lists(n, rate, alpha, ft, spot, ccd):
    cols = _synthetic(n, rate, 1234 + n, spot=spot, ccd=ccd)
    args = ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"]
    ref = P.oracle_pileup(cols, args, "c1_acis_s", 77)
    assert len(ref["t"]) > 0 and (ref["nphotons"] >= 2).any()
    with marx_b200.MarxB200(PACK["pileup_acis_s_bright"], seed=1, max_photons=1024) as m:
        before = m.launch_count()
        got, _ = m.pileup(cols, alpha, ft, 77)
        launches = m.launch_count() - before
    # frames of up to 513 events stay in the fused kernel; the 10^4-events-per-frame list makes it hand over to the step kernels
    assert launches == (12 if rate * ft > 600 else 4), launches
    _same(got, ref, "synthetic")
    assert int(got["nphotons"].sum()) <= n and (np.diff(got["frame"]) >= 0).all()


def test_edges_and_errors():
    cols, ref, seed = P.load_golden("pileup_acis_s_moderate")
    alpha, ft = P.pileup_params(P.CASES["pileup_acis_s_moderate"][1])
    with marx_b200.MarxB200(PACK["pileup_acis_s_moderate"], seed=1, max_photons=1024) as m:
        empty, _ = m.pileup({k: v[:0] for k, v in cols.items()}, alpha, ft, seed)
        assert all(len(v) == 0 for v in empty.values())
        one, _ = m.pileup({k: v[:1] for k, v in cols.items()}, alpha, ft, seed)
        assert len(one["t"]) == 1 and one["nphotons"][0] == 1 and one["x"][0] == cols["x"][0] and one["benergy"][0] == cols["benergy"][0]
        # identical events in one frame pile into one island; alpha = 1 keeps it
        same = {k: np.repeat(v[:1], 5) for k, v in cols.items()}
        piled, _ = m.pileup(same, 1.0, ft, seed)
        assert len(piled["t"]) == 1 and piled["nphotons"][0] == 5
        assert np.isclose(piled["benergy"][0], 5.0 * cols["benergy"][0], rtol=1e-6)
        with pytest.raises(marx_b200.MarxB200Error, match="max_out"):
            m.pileup(cols, alpha, ft, seed, max_out=len(ref["t"]) - 1)
        exact, _ = m.pileup(cols, alpha, ft, seed, max_out=len(ref["t"]))
        _same(exact, ref, "exact capacity")
        bad = dict(cols, ccd=np.full(len(cols["t"]), 11, np.int8))
        with pytest.raises(marx_b200.MarxB200Error, match="CCD"):
            m.pileup(bad, alpha, ft, seed)
        bad = dict(cols, x=np.full(len(cols["t"]), 2000.0, np.float32))
        with pytest.raises(marx_b200.MarxB200Error, match="corrupt"):
            m.pileup(bad, alpha, ft, seed)
        long_frame = {k: np.repeat(v[:1], 70000) for k, v in cols.items()}
        with pytest.raises(marx_b200.MarxB200Error, match="65536"):
            m.pileup(long_frame, alpha, ft, seed)
        # the context still works after the refusals
        again, _ = m.pileup(cols, alpha, ft, seed)
        _same(again, ref, "after errors")
    with marx_b200.MarxB200(marx_b200.caldata_path("c3_letg_hrc_s.calpack"), seed=1, max_photons=1024) as m:
        with pytest.raises(marx_b200.MarxB200Error, match="ACIS"):
            m.pileup(cols, alpha, ft, seed)


def test_pileup_of_the_device_resident_event_list():
    """marxb200_pileup_events: `marx` then `marxpileup` without the column files in between.  The rows must equal
    marxb200_pileup_run on the columns marx_write_photons would have written for the same list (and so the oracle's)."""
    n, total_time, alpha, ft = 1 << 21, 1234.5, 0.5, 3.241
    with marx_b200.MarxB200(PACK["pileup_acis_s_bright"], seed=21, max_photons=n) as m:
        m.trace(0, n, time_base=0.0)
        ph = m.download().copy()
        dev, ms = m.pileup_events(total_time, alpha, ft, 5)
        cols = {"ccd": ph["ccd_num"].astype(np.int8), "x": ph["y_pixel"].astype(np.float32), "y": ph["z_pixel"].astype(np.float32),
                "t": (ph["arrival_time"] + total_time).astype(np.float32), "benergy": ph["pi"].astype(np.float32),
                **{k: np.ascontiguousarray(ph["dither"][:, j]).astype(np.float32) for j, k in enumerate(P.DITHER)}}
        host, _ = m.pileup(cols, alpha, ft, 5)
        assert ms > 0.0 and len(dev["t"]) > 0 and (dev["nphotons"] >= 2).any()
        _same(dev, host, "device-resident list")
        ref = P.oracle_pileup(cols, ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"], "c1_acis_s", 5)
        _same(dev, ref, "device-resident list vs oracle")
