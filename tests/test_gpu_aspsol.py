"""DitherModel=FILE (ASPSOL) on the GPU beyond the slot-by-slot parameter-surface cases: the compacting trace path over
several batches with the running clock, the end of the aspect solution inside / before a batch, the detector-dither
columns of the event files, and the unmodified marx driver end to end (dither.c:288-500, detector.c:275-295)."""
import os
import re
import subprocess

import numpy as np
import pytest

import marx_b200
from tests.golden.make_golden import COMMON, write_aspsol_fits
from tests.test_gpu_marx_driver import MARX_GPU, needs_driver, read_dir, run_marx
from tests.test_oracle_vs_reference import HAVE_REF, REF

pytestmark = pytest.mark.gpu
ARGS = ["SourceType=POINT", "MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=FILE"]


def make_pack(tmp_path, duration):
    asol = write_aspsol_fits(str(tmp_path / "asol1.fits"), duration=duration)
    pack = str(tmp_path / "aspsol.calpack")
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    subprocess.check_call([os.path.join(REF, "calpack_dump"), pack, "@@" + os.path.join(REF, "par", "marx.par")] + COMMON + ARGS
                          + ["DitherFile=" + asol], env=env, stdout=subprocess.DEVNULL)
    return pack, asol


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
def test_batched_trace_with_running_clock_and_end_of_file(tmp_path):
    """four batches of 2^15 rays (~4.8 ks each) against a 12 ks aspect solution (the file ends inside the third batch): the compacted
    event lists equal the oracle's survivors, the third batch is cut where the oracle's is, the fourth is empty"""
    from tests.oracle_lib import Oracle
    from tests.parity import assert_stage_ok, compare_stage
    pack, _ = make_pack(tmp_path, 12000.0)
    n, seed = 1 << 15, 61
    o = Oracle(pack, seed)
    tb, kept_total = 0.0, []
    with marx_b200.MarxB200(pack, seed=seed, max_photons=n) as m:
        for b in range(4):
            ref, tb_new, n_det = o.trace(b * n, n, time_base=tb)
            kept = o.last_generated
            m.trace(b * n, n, time_base=(0.0 if b == 0 else -1.0))
            gen, live, t_end = m.counts()
            assert gen == kept and live == n_det, (b, gen, kept, live, n_det)
            kept_total.append(kept)
            if kept:
                assert abs(t_end - tb_new) <= 1e-12 * tb_new
                got = m.download()
                want = ref[3][:kept]
                want = want[(want["flags"] & 0xFF) == 0]
                assert (got["tag"] == want["tag"]).all()
                # a last-bit difference of the (parallel) time sum can flip the float rounding of an interpolated angle
                # (see check_cuda_against_oracle): those rays are compared at 1e-7, the others at 1e-9
                flipped = (got["dither"] != want["dither"]).any(axis=1)
                assert flipped.mean() <= 2e-3
                assert_stage_ok(compare_stage(got[~flipped], want[~flipped], 3), 3)
                if flipped.any():
                    g = compare_stage(got[flipped], want[flipped], 3)
                    assert g["p_max_abs"] <= 1e-7 and g.get("x_max_rel", 0) <= 1e-7 and g.get("pha_mismatch", 0) == 0, g
                assert np.abs(got["dither"] - want["dither"]).max() < 1e-9
                assert (got["dither"][:, 3:] != 0).any()
                cols = m.download_columns(("ray", "chipx", "chipy", "pha"))
                assert (cols["ray"] == want["tag"]).all() and (cols["pha"] == want["pulse_height"]).all()
            tb = tb_new
    assert kept_total[0] == n and kept_total[1] == n and 0 < kept_total[2] < n and kept_total[3] == 0, kept_total


@needs_driver
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
def test_marx_driver_with_aspsol_file(tmp_path):
    """the unmodified driver: DitherModel=FILE output equals the C ABI's events (sky and detector dither columns
    included), and the run stops where the aspect solution ends, as the stock CPU marx does"""
    pack, asol = make_pack(tmp_path, 9000.0)
    dn, seed = 20000, 7
    args = COMMON + ARGS + ["DitherFile=" + asol, "NumRays=200000", "dNumRays=%d" % dn, "RandomSeed=%d" % seed, "Verbose=1"]
    p = run_marx(MARX_GPU, tmp_path / "out", args)
    assert "marxb200: ray trace on CUDA device" in p.stdout
    got = read_dir(tmp_path / "out")
    want, total = [], 0
    with marx_b200.MarxB200(pack, seed=seed, max_photons=dn) as m:
        first = 0
        while True:
            m.trace(first, dn, time_base=(0.0 if first == 0 else -1.0))
            gen, live, _ = m.counts()
            total += gen
            if gen:
                want.append(m.download().copy())
            if gen < dn:
                break
            first += dn
    ph = np.concatenate(want)
    assert 0 < total < 200000
    tot, det = re.findall(r"Total photons: (\d+), Total Photons detected: (\d+)", p.stdout)[-1]
    assert int(tot) == total and int(det) == len(ph) == len(got["energy.dat"])
    assert (got["tag.dat"].astype(np.uint32) == ph["tag"]).all()
    assert (got["pha.dat"] == ph["pulse_height"]).all() and (got["order.dat"] == ph["order"]).all()
    assert (got["xpixel.dat"] == ph["y_pixel"]).all() and (got["ypixel.dat"] == ph["z_pixel"]).all()
    for k, f in enumerate(("sky_ra.dat", "sky_dec.dat", "sky_roll.dat", "det_dy.dat", "det_dz.dat", "det_theta.dat")):
        assert (got[f] == ph["dither"][:, k]).all(), f
    assert (got["det_dy.dat"] != 0).any() and (got["det_theta.dat"] != 0).any()
    # the stock CPU marx (own RNG) ends at the same place: the same exposure, rays within Poisson noise
    q = run_marx(os.path.join(REF, "marx"), tmp_path / "cpu", args)
    tot_cpu, det_cpu = re.findall(r"Total photons: (\d+), Total Photons detected: (\d+)", q.stdout)[-1]
    assert abs(int(tot_cpu) - total) < 6 * np.sqrt(total)
    assert abs(int(det_cpu) - len(ph)) < 6 * np.sqrt(len(ph))
