"""Pins the CPU restatement (oracle/marx_oracle.c) to the reference itself.

1. Against the committed fixtures: per-stage FP64 photon records produced by the UNMODIFIED MARX 5.5.3
   stage functions (oracle/_ref/marx_replay, see tests/golden/make_golden.py) -- required BIT-EXACT.
2. Where oracle/_ref exists (the build container), against freshly generated replays with other seeds,
   ray ranges and parameter overrides -- also bit-exact.
"""
import os
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import Oracle
from tests.replay_io import read_replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = os.path.exists(os.path.join(REF, "marx_replay")) and os.path.exists(os.path.join(REF, "calpack_dump"))

LIVE_FIELDS = ["energy", "x", "p", "flags", "y_pixel", "z_pixel", "u_pixel", "v_pixel", "dither", "pi", "pulse_height",
               "mirror_shell", "ccd_num", "detector_region", "order", "support_orders", "tag"]


def check_bit_exact(mine, ref_stages, start_times):
    """mine: oracle records [4][n] with absolute times; ref_stages: replay records [n][4]."""
    n = mine.shape[1]
    for s in range(4):
        a, b = mine[s], ref_stages[:, s]
        a_alive = (a["flags"] & 0xFF) == 0
        b_alive = (b["flags"] & 0xFF) == 0
        assert (a_alive == b_alive).all(), "stage %d: live sets differ" % s
        # dead rays: first cause reported by the restatement must be among the reference's bits
        dead = ~a_alive
        assert ((a["flags"][dead] & b["flags"][dead] & 0xFF) == (a["flags"][dead] & 0xFF)).all()
        for f in LIVE_FIELDS:
            if s == 0 and f == "x":
                continue
            assert (a[f][a_alive] == b[f][a_alive]).all(), "stage %d field %s differs" % (s, f)
    # absolute arrival times: reference (batch size 1) = running start_time + arrival_time
    t_ref = start_times + ref_stages[:, 0]["arrival_time"]
    assert np.abs(mine[0]["arrival_time"] - t_ref).max() <= 1e-9 * max(t_ref.max(), 1.0)
    assert n == len(ref_stages)


@pytest.mark.parametrize("config", ["c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s", "c4_beta_acis_i", "c4_image_acis_i", "c1_line_acis_s", "c3_hrc_i"])
def test_oracle_matches_committed_reference_replay(config):
    z = np.load(os.path.join(GOLDEN, config + "_replay.npz"))
    o = Oracle(config, int(z["seed"]))
    st, _, nd = o.trace(int(z["first_ray"]), len(z["stages"]))
    check_bit_exact(st, z["stages"], z["start_time"])
    assert nd == int(((z["stages"][:, 3]["flags"] & 0xFF) == 0).sum())


CASES = [
    ("hetg_mid_stream", ["MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL"], 3, 5000000, 30000),
    ("soft_no_grating_dither", ["MinEnergy=0.1", "MaxEnergy=1.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL"], 9, 0, 20000),
    ("hard_hetg_nodither", ["MinEnergy=5.0", "MaxEnergy=11.5", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=NONE"], 21, 123456, 20000),
    ("ideal_mirror", ["MinEnergy=1.0", "MaxEnergy=2.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "HRMA_Ideal=yes"], 4, 0, 10000),
    ("no_wfold_no_struts_offset_detector", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                                             "HRMA_Use_WFold=no", "HRMA_Use_Struts=no", "DetOffsetX=0.5", "DetOffsetZ=-3.0"], 5, 77, 10000),
    ("finite_distance_shutters", ["MinEnergy=0.5", "MaxEnergy=4.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=NONE",
                                  "SourceDistance=537.0", "Shutters1=0110", "Shutters4=1000"], 6, 0, 10000),
    ("gauss_source_acis_i", ["SourceType=GAUSS", "S-GaussSigma=45", "MinEnergy=0.5", "MaxEnergy=5.0", "GratingType=NONE",
                             "DetectorType=ACIS-I", "DitherModel=INTERNAL"], 15, 0, 10000),
    ("disk_source_hetg", ["SourceType=DISK", "S-DiskTheta0=20", "S-DiskTheta1=90", "MinEnergy=0.8", "MaxEnergy=2.5", "GratingType=HETG",
                          "DetectorType=ACIS-S", "DitherModel=NONE"], 16, 4096, 10000),
    ("beta_source_off_axis_acis_i", ["SourceType=BETA", "S-BetaCoreRadius=25", "S-BetaBeta=0.9", "SourceRA=249.9316", "MinEnergy=0.5",
                                     "MaxEnergy=7.0", "GratingType=NONE", "DetectorType=ACIS-I", "DitherModel=INTERNAL"], 17, 0, 10000),
    ("file_spectrum", ["SpectrumType=FILE", "SpectrumFile=%SPECFILE%", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL"], 18, 0, 10000),
    ("letg_hrc_s_no_hesf", ["MinEnergy=0.08", "MaxEnergy=3.0", "GratingType=LETG", "DetectorType=HRC-S", "DitherModel=INTERNAL",
                            "HRC-HESF=no"], 19, 0, 20000),
    ("hetg_hrc_s_hesf", ["MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=HRC-S", "DitherModel=NONE"], 20, 0, 20000),
    ("letg_acis_s", ["MinEnergy=0.3", "MaxEnergy=4.0", "GratingType=LETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL"], 22, 0, 10000),
    ("letg_computed_efficiencies", ["MinEnergy=0.2", "MaxEnergy=2.0", "GratingType=LETG", "DetectorType=HRC-S", "DitherModel=INTERNAL",
                                    "UseGratingEffFiles=no"], 23, 0, 10000),
    ("det_ideal", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "DetIdeal=yes"], 31, 0, 8000),
    ("no_blur_vignetting", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                            "HRMA_Use_Blur=no", "HRMAVig=0.8"], 32, 0, 8000),
    ("no_scale_factors_big_blurs", ["MinEnergy=1.0", "MaxEnergy=7.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                                    "HRMA_Use_Scale_Factors=no", "AspectBlur=1.5", "P1Blur=0.6", "H6Blur=0.2"], 33, 65536, 8000),
    ("roll_dither_off_axis", ["MinEnergy=0.5", "MaxEnergy=4.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                              "DitherAmp_Roll=30", "DitherPeriod_Roll=700", "SourceRA=250.05", "Roll_Nom=17.0"], 34, 0, 8000),
    ("hrc_i_letg", ["MinEnergy=0.1", "MaxEnergy=1.5", "GratingType=LETG", "DetectorType=HRC-I", "DitherModel=INTERNAL"], 37, 0, 8000),
    ("detector_none", ["MinEnergy=0.5", "MaxEnergy=4.0", "GratingType=HETG", "DetectorType=NONE", "DitherModel=NONE"], 38, 0, 8000),
    ("det_extend_acis_s_off_axis", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                                    "DetExtendFlag=yes", "SourceDEC=-53.45"], 41, 0, 12000),
    ("det_extend_hrc_s_letg", ["MinEnergy=0.07", "MaxEnergy=1.0", "GratingType=LETG", "DetectorType=HRC-S", "DitherModel=INTERNAL",
                               "DetExtendFlag=yes", "SourceRA=250.6"], 42, 0, 12000),
    ("det_extend_acis_i", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=NONE", "DetectorType=ACIS-I", "DitherModel=NONE",
                           "DetExtendFlag=yes", "SourceRA=250.35"], 43, 0, 12000),
    # DitherModel=FILE (dither.c:288-500): synthetic aspect solution with pointing, roll AND SIM motion,
    # so the per-photon detector dither (detector.c:275-295) is live; the third case ends the file inside the run
    ("aspsol_hetg_acis_s", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=FILE",
                            "DitherFile=%ASPSOL:6000%"], 51, 0, 12000),
    ("aspsol_letg_hrc_s", ["MinEnergy=0.1", "MaxEnergy=2.0", "GratingType=LETG", "DetectorType=HRC-S", "DitherModel=FILE",
                           "DitherFile=%ASPSOL:6000%", "AspectBlur=0.2"], 52, 0, 12000),
    ("aspsol_ends_early_acis_i", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=NONE", "DetectorType=ACIS-I", "DitherModel=FILE",
                                  "DitherFile=%ASPSOL:1500%"], 53, 0, 12000),
    # 64-bit ray indices: beyond 2^32 the record's tag (marx.h:98) wraps and only the full index keys the draws; the second
    # batch straddles the boundary
    ("ray_index_beyond_2_32", ["MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL"],
     3, (1 << 33) + 12345, 12000),
    ("ray_index_across_2_32", ["MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL"],
     3, (1 << 32) - 6000, 12000),
    # MirrorType=FLATFIELD (ffield.c): no optics, rays start on a rectangle in front of the detector (24 cm^2 instead of the HRMA's
    # 1145: the flux is raised so that arrival times stay in the range of the other cases)
    ("flatfield_acis_s", ["MirrorType=FLATFIELD", "FF_MinY=-30", "FF_MaxY=45", "FF_MinZ=-20", "FF_MaxZ=12", "FF_XPos=9000", "MinEnergy=0.5",
                          "MaxEnergy=7.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "SourceFlux=0.15"], 61, 0, 10000),
    ("flatfield_finite_distance_hetg_hrc_s", ["MirrorType=FLATFIELD", "SourceDistance=537.0", "MinEnergy=0.5", "MaxEnergy=3.0",
                                              "GratingType=HETG", "DetectorType=HRC-S", "DitherModel=NONE"], 62, 4096, 10000),
    ("sector_files_off_unit_order", ["MinEnergy=0.8", "MaxEnergy=3.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                                     "Use_HETG_Sector_Files=no"], 8, 1000, 10000),
]


def expand_args(tmp_path, args):
    """materialise the files a case refers to (%SPECFILE%, %IMAGE%, %ASPSOL:<seconds>%) and default the source type"""
    from tests.golden.make_golden import write_aspsol_fits, write_beta_image_fits
    if any("%SPECFILE%" in a for a in args):
        # a FILE spectrum (spectrum.c:214-273): two columns, energy [keV] and flux density
        spec = tmp_path / "spec.dat"
        e = np.linspace(0.4, 9.0, 400)
        spec.write_text("".join("%.6f %.6e\n" % (x, x ** -1.7 * (1 + 3 * np.exp(-0.5 * ((x - 6.4) / 0.05) ** 2))) for x in e))
        args = [a.replace("%SPECFILE%", str(spec)) for a in args]
    if any("%IMAGE%" in a for a in args):
        args = [a.replace("%IMAGE%", write_beta_image_fits(str(tmp_path / "img.fits"), n=128, cdelt_arcsec=1.0)) for a in args]
    out = []
    for a in args:
        if "%ASPSOL:" in a:
            dur = float(a.split("%ASPSOL:")[1].rstrip("%"))
            a = a.split("%ASPSOL:")[0] + write_aspsol_fits(str(tmp_path / "asol1.fits"), duration=dur)
        out.append(a)
    if not any(a.startswith("SourceType=") for a in out):
        out = ["SourceType=POINT"] + out
    return out


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
@pytest.mark.parametrize("name,args,seed,first,n", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_fresh_reference_replay(tmp_path, name, args, seed, first, n):
    par = "@@" + os.path.join(REF, "par", "marx.par")
    common = ["ExposureTime=0", "Verbose=0", "SourceFlux=0.003", "TStart=2023.5", "SpectrumType=FLAT"]
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    args = expand_args(tmp_path, args)
    pack = str(tmp_path / (name + ".calpack"))
    dump = str(tmp_path / (name + ".bin"))
    subprocess.check_call([os.path.join(REF, "calpack_dump"), pack, par] + common + args, env=env,
                          stdout=subprocess.DEVNULL)
    subprocess.check_call([os.path.join(REF, "marx_replay"), dump, str(n), str(seed), str(first), par] + common + args,
                          env=env, stdout=subprocess.DEVNULL)
    hdr, recs = read_replay(dump)
    o = Oracle(pack, seed)
    st, _, _ = o.trace(first, n)
    kept = o.last_generated
    if "ends_early" in name:
        assert 0 < kept < n, "the ASPSOL file was meant to end inside the run"
        assert (st[:, kept:].view(np.uint8) == 0).all()
    else:
        assert kept == n
    assert kept == hdr["nrays"] == len(recs)
    check_bit_exact(st[:, :kept], recs["st"], recs["start"])
