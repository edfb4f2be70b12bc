"""GPU parity of the Level-1 event transforms (marxb200_level1_*), through the C ABI:
 * the events of the committed marx2fits fixtures injected into the device list -> integer columns bit-exact against the stock
   marx2fits table, float32 columns within 1 ulp (device libm), FP64 columns within 1e-9 relative of the oracle;
 * the device's own traced events (C2, 2^22 rays, two consecutive batches of one event file) against the oracle fed with the
   same column values -- the exposure-frame aspect broadcast crossing tile and batch boundaries."""
import os

import numpy as np
import pytest

import marx_b200
from marx_b200.level1 import Level1Desc
from tests import level1_lib as L

pytestmark = pytest.mark.gpu

INT_COLS = ["expno", "tdetx", "tdety", "pha", "hrc_u", "hrc_v", "ccd_id", "node_id", "chipx", "chipy", "pi", "fltgrade", "grade", "status", "keep"]
F64_COLS = ["time", "detx", "dety", "x", "y"]
CALPACK = {"level1_acis_s_hetg_edser": "c2_hetg_acis_s", "level1_acis_i_beta_randomize": "c4_beta_acis_i", "level1_acis_s_nodither_none": "c1_acis_s",
           "level1_acis_s_hetg_exact": "c2_hetg_acis_s", "level1_hrc_s_letg": "c3_letg_hrc_s"}


def assert_matches_oracle(gpu, ora, desc):
    for k in INT_COLS:
        assert np.array_equal(gpu[k], ora[k]), (k, np.flatnonzero(gpu[k] != ora[k])[:5])
    assert np.array_equal(gpu["energy"], ora["energy"])
    for k in ("time", "detx", "dety"):
        # 1e-9 relative (north_star); measured ~1e-15: only the device's sin/cos and its reciprocal normalisation differ from glibc
        assert np.allclose(gpu[k], ora[k], rtol=1e-9, atol=0), (k, np.abs(gpu[k] / ora[k] - 1).max())
    assert np.allclose(gpu["x"], ora["x"], rtol=1e-9, atol=0), np.abs(gpu["x"] / ora["x"] - 1).max()
    if int(desc["used_dither"]) == 0:
        assert np.allclose(gpu["y"], ora["y"], rtol=1e-9, atol=0)
    else:
        # Y alone goes through the reference's ill-conditioned acos (tests/level1_lib.py: sky_y_tolerance): 4 ulps of its argument
        assert (np.abs(gpu["y"] - ora["y"]) <= L.sky_y_tolerance(ora["y"], desc)).all(), np.abs(gpu["y"] - ora["y"]).max()


@pytest.mark.parametrize("case", sorted(L.LEVEL1_CASES))
def test_level1_fixture_events_match_the_stock_marx2fits_table(case):
    desc, cols, ref, seed = L.load_golden(case)
    n = len(cols["time"])
    with marx_b200.MarxB200(CALPACK[case], seed=seed, max_photons=n + 1024) as m:
        m.set_level1(Level1Desc.from_dict(desc))
        m.upload(L.photons_from_columns(cols), start_time=0.0)
        m.level1_transform(0.0)
        gpu = m.level1_download()
    assert len(gpu["time"]) == n
    ora = L.Level1Oracle(desc, seed).transform(cols)
    ytol = L.sky_y_tolerance(ora["y"], desc) if int(desc["used_dither"]) else None
    L.compare_with_fits(gpu, ref, f32_ulps=1, y_tolerance=ytol)
    assert_matches_oracle(gpu, ora, desc)


def test_level1_fixture_in_pieces_carries_the_exposure_state():
    desc, cols, ref, seed = L.load_golden("level1_acis_s_hetg_edser")
    n = len(cols["time"])
    ph = L.photons_from_columns(cols)
    cuts = [0, 1, 300, 301, 2000, n]
    parts = []
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=seed, max_photons=n + 1024) as m:
        m.set_level1(Level1Desc.from_dict(desc))
        for a, b in zip(cuts[:-1], cuts[1:]):
            m.upload(ph[a:b], start_time=0.0)
            m.level1_transform(0.0)
            parts.append(m.level1_download())
        gpu = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        ora = L.Level1Oracle(desc, seed).transform(cols)
        L.compare_with_fits(gpu, ref, f32_ulps=1, y_tolerance=L.sky_y_tolerance(ora["y"], desc))
        # a new file starts from scratch again
        m.level1_reset()
        m.upload(ph, start_time=0.0)
        m.level1_transform(0.0)
        again = m.level1_download()
    for k in INT_COLS:
        assert np.array_equal(again[k], gpu[k]), k


@pytest.mark.parametrize("pixadj", ["edser", "randomize"])
def test_level1_of_traced_events_matches_the_oracle(pixadj):
    desc, _, _, _ = L.load_golden("level1_acis_s_hetg_edser")
    desc = dict(desc, pix_adjust=marx_b200.level1.PIXADJ[pixadj])
    n, seed = 1 << 22, 5
    names = ("time", "chipx", "chipy", "pi", "pha", "ccd")
    got, inputs = [], []
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=seed, max_photons=n) as m:
        m.set_level1(Level1Desc.from_dict(desc))
        total_time = 0.0
        for b in range(2):
            m.trace(b * n, n)
            m.level1_transform(total_time)
            got.append(m.level1_download())
            ph = m.download()
            # the values marx_write_photons would put into the column files (marxio.c:217-290)
            inputs.append({"time": (ph["arrival_time"] + total_time).astype(np.float32), "xpixel": ph["y_pixel"], "ypixel": ph["z_pixel"],
                           "b_energy": ph["pi"], "pha": ph["pulse_height"], "ccd": ph["ccd_num"],
                           **{k: np.ascontiguousarray(ph["dither"][:, j]) for j, k in enumerate(L.DITHER_KEYS)}})
            total_time = m.counts()[2]
    assert len(got[0]["time"]) > 250000
    o = L.Level1Oracle(desc, seed)
    for g, cols in zip(got, inputs):
        ora = o.transform(cols)
        assert_matches_oracle(g, ora, desc)
    # exposure frames hold several events: the broadcast is exercised (and crosses 256-event tiles)
    e = got[0]["expno"]
    assert (e[1:] == e[:-1]).mean() > 0.3


MARX2FITS_GPU = os.path.join(L.ROOT, "integration", "_build", "marx2fits_gpu")


@pytest.mark.skipif(not (L.have_reference() and os.path.exists(MARX2FITS_GPU)), reason="oracle/_ref or integration/_build not built")
@pytest.mark.parametrize("case", ["level1_acis_s_hetg_edser", "level1_acis_i_beta_randomize", "level1_hrc_s_letg"])
def test_marx2fits_gpu_writes_the_stock_events_table(case, tmp_path):
    """integration/marx2fits_gpu = the unmodified marx2fits.c (header, readers, jdfits table writer) with its computed columns
    re-pointed at marxb200_level1_*: its EVENTS table against the stock program's (same per-row draw stream) for a fresh stock
    marx run -- integers exact, float32 columns within one ulp, Y within the reference formula's conditioning bound."""
    import subprocess
    args, pixadj, ndraw = L.LEVEL1_CASES[case]
    out = tmp_path / "out"
    L.run_stock_marx(out, args, n_rays=300000, seed=9)
    ref = L.run_stock_marx2fits(out, tmp_path / "ref.fits", pixadj, ndraw, seed=21)
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(L.REF, "data"), L1_SEED="21")
    p = subprocess.run([MARX2FITS_GPU, "--pixadj=" + pixadj, str(out), str(tmp_path / "gpu.fits")], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    from tests.fits_table import read_bintable
    gpu, hdr = read_bintable(str(tmp_path / "gpu.fits"), "EVENTS")
    assert set(gpu) == set(read_bintable(str(tmp_path / "ref.fits"), "EVENTS")[0])
    desc = L.dump_descriptor(out, pixadj)
    ytol = None
    if int(desc["used_dither"]):
        ytol = L.sky_y_tolerance(ref["Y"].astype(np.float64), desc) + 2.0 * np.spacing(np.abs(ref["Y"]).astype(np.float32)).astype(np.float64)
    for name, r in ref.items():
        g = gpu[name]
        if name == "STATUS":
            g = (g.astype(np.uint32) << np.array([24, 16, 8, 0], dtype=np.uint32)).sum(axis=1).astype(np.int64)
        assert len(g) == len(r) > 5000, name
        if name == "Y" and ytol is not None:
            assert (np.abs(g.astype(np.float64) - r.astype(np.float64)) <= ytol).all(), name
        elif r.dtype.kind == "f" and r.dtype.itemsize == 4:
            d = np.abs(g.view(np.int32).astype(np.int64) - r.view(np.int32).astype(np.int64))
            assert d.max() <= 1, (name, int(d.max()))
        elif r.dtype.kind == "f":
            assert np.allclose(g, r, rtol=1e-12, atol=0), name
        else:
            assert np.array_equal(g.astype(np.int64), r.astype(np.int64)), name
