#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference built in oracle/_ref.

Run in the build container (needs oracle/_ref, i.e. `make -C oracle/ref`):
    python tests/golden/make_golden.py
For every config it (1) writes the calibration pack the CUDA path loads (marx_b200/caldata/*.calpack,
via oracle/_ref/calpack_dump = the stock *_init functions) and (2) runs oracle/_ref/marx_replay (stock
stage functions, Philox draws, batch size 1) and stores the per-stage FP64 photon records as
tests/golden/<config>_replay.npz.  Nothing here runs on the GPU box.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
from tests.replay_io import read_replay  # noqa: E402

IMAGE_FITS = "/tmp/marxb200_beta_image.fits"


def write_beta_image_fits(path=IMAGE_FITS, n=256, cdelt_arcsec=0.5, core_arcsec=10.0, beta=0.7):
    """synthetic IMAGE source: an n x n float32 beta-model surface-brightness map as a primary
    FITS HDU with CDELT1/2, the keywords s-image.c:200-262 reads."""
    yy, xx = np.mgrid[0:n, 0:n]
    r2 = ((xx - n / 2) ** 2 + (yy - n / 2) ** 2) * cdelt_arcsec ** 2
    img = (1.0 + r2 / core_arcsec ** 2) ** (-3.0 * beta + 0.5)
    cards = [("SIMPLE", "T"), ("BITPIX", "-32"), ("NAXIS", "2"), ("NAXIS1", str(n)), ("NAXIS2", str(n)),
             ("CDELT1", "%.12E" % (-cdelt_arcsec / 3600.0)), ("CDELT2", "%.12E" % (cdelt_arcsec / 3600.0))]
    hdr = "".join(("%-8s= %20s" % kv).ljust(80) for kv in cards) + "END".ljust(80)
    hdr = hdr.ljust((len(hdr) + 2879) // 2880 * 2880)
    data = img.astype(">f4").tobytes()
    data += b"\0" * (-len(data) % 2880)
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii") + data)
    return path


def write_aspsol_fits(path, tstart=7.9e8, duration=6000.0, step=2.05, ra_nom=250.2134679741175, dec_nom=-53.75743813458669,
                      roll_nom=237.36968458476):
    """synthetic aspect solution for DitherModel=FILE: one BINTABLE 'ASPSOL' with the seven double columns and the four
    keywords init_aspsol_dither reads (dither.c:432-500); angles in degrees, dy/dz in mm, like a CXC asol1 file.
    Lissajous pointing dither (16 arcsec), a slow roll drift and a small SIM motion (dy, dz, dtheta all non-zero, so the
    per-photon detector dither of detector.c:275-295 is exercised)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref"))
    from synth_acis_caldb import bintable_hdu, primary_hdu
    n = int(duration / step) + 1
    t = tstart + step * np.arange(n)
    ph = 2 * np.pi * (t - tstart)
    asec = 1.0 / 3600.0
    cols = [
        ("time", "D", 1, t),
        ("ra", "D", 1, ra_nom + 16 * asec / np.cos(np.radians(dec_nom)) * np.sin(ph / 1000.0)),
        ("dec", "D", 1, dec_nom + 16 * asec * np.sin(ph / 707.1 + 0.3)),
        ("roll", "D", 1, roll_nom + 25 * asec * np.sin(ph / 3100.0 + 1.1)),
        ("dy", "D", 1, 0.040 * np.sin(ph / 1800.0) + 0.003),
        ("dz", "D", 1, 0.025 * np.cos(ph / 2600.0) - 0.002),
        ("dtheta", "D", 1, 2.0e-3 * np.sin(ph / 4100.0 + 0.7)),
    ]
    keys = [("RA_NOM", float(ra_nom)), ("DEC_NOM", float(dec_nom)), ("ROLL_NOM", float(roll_nom)), ("TSTART", float(tstart))]
    with open(path, "wb") as f:
        f.write(primary_hdu() + bintable_hdu("ASPSOL", cols, n, extra_keys=keys))
    return path


COMMON = ["ExposureTime=0", "Verbose=0", "SourceFlux=0.003", "TStart=2023.5", "SpectrumType=FLAT"]
CONFIGS = {
    # BASELINE.json configs[0]
    "c1_acis_s": dict(args=["SourceType=POINT", "MinEnergy=1.5", "MaxEnergy=1.5", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=NONE"],
                      nrays=8192, seed=11),
    # BASELINE.json configs[1] (the bench workload)
    "c2_hetg_acis_s": dict(args=["SourceType=POINT", "MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S",
                                 "DitherModel=INTERNAL"], nrays=16384, seed=7),
    # BASELINE.json configs[2]: LETG + HRC-S (HESF on, fine + coarse support gratings)
    "c3_letg_hrc_s": dict(args=["SourceType=POINT", "MinEnergy=0.1", "MaxEnergy=2.0", "GratingType=LETG", "DetectorType=HRC-S",
                                "DitherModel=INTERNAL"], nrays=16384, seed=9),
    # BASELINE.json configs[3]: extended BETA source 10 arcmin off axis, ACIS-I, dither
    "c4_beta_acis_i": dict(args=["SourceType=BETA", "S-BetaCoreRadius=10", "S-BetaBeta=0.7", "SourceDEC=-53.92410480125",
                                 "MinEnergy=0.5", "MaxEnergy=7.0", "GratingType=NONE", "DetectorType=ACIS-I",
                                 "DitherModel=INTERNAL"], nrays=16384, seed=13),
    # the IMAGE flavour of configs[3]: synthetic 256^2 beta-model FITS image (0.5 arcsec pixels), same pointing offset
    "c4_image_acis_i": dict(args=["SourceType=IMAGE", "S-ImageFile=" + IMAGE_FITS, "SourceDEC=-53.92410480125",
                                  "MinEnergy=0.5", "MaxEnergy=7.0", "GratingType=NONE", "DetectorType=ACIS-I",
                                  "DitherModel=INTERNAL"], nrays=16384, seed=14),
    # HRC-I (hrc-i.c), no grating
    "c3_hrc_i": dict(args=["SourceType=POINT", "MinEnergy=0.1", "MaxEnergy=2.0", "GratingType=NONE", "DetectorType=HRC-I",
                           "DitherModel=INTERNAL"], nrays=8192, seed=16),
    # LINE source (s-line.c), no grating, ACIS-S, no dither
    "c1_line_acis_s": dict(args=["SourceType=LINE", "S-LinePhi=30", "S-LineTheta=60", "MinEnergy=1.0", "MaxEnergy=2.0",
                                 "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=NONE"], nrays=8192, seed=15),
}


def main():
    par = "@@" + os.path.join(REF, "par", "marx.par")
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    write_beta_image_fits()
    only = sys.argv[1:]
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        pack = os.path.join(ROOT, "marx_b200", "caldata", name + ".calpack")
        subprocess.check_call([os.path.join(REF, "calpack_dump"), pack, par] + COMMON + cfg["args"], env=env)
        tmp = os.path.join("/tmp", name + "_replay.bin")
        subprocess.check_call([os.path.join(REF, "marx_replay"), tmp, str(cfg["nrays"]), str(cfg["seed"]), "0", par]
                              + COMMON + cfg["args"], env=env)
        hdr, recs = read_replay(tmp)
        out = os.path.join(ROOT, "tests", "golden", name + "_replay.npz")
        # a dead ray's record is only meaningful through its flags: blank the rest so the fixture stays small
        recs = recs.copy()
        for s in range(4):
            st = recs["st"][:, s]
            dead = (st["flags"] & 0xFF) != 0
            keep_flags, keep_tag = st["flags"][dead].copy(), st["tag"][dead].copy()
            st[dead] = np.zeros(1, dtype=st.dtype)
            st["flags"][dead] = keep_flags
            st["tag"][dead] = keep_tag
            recs["st"][:, s] = st
        np.savez_compressed(out, seed=hdr["seed"], first_ray=hdr["first_ray"], modules=hdr["modules"],
                            stages=recs["st"], draws=recs["draws"], start_time=recs["start"])
        alive = [(recs["st"][:, s]["flags"] & 0xFF == 0).sum() for s in range(4)]
        print(name, "rays", len(recs), "alive per stage", alive, "->", out, os.path.getsize(out) // 1024, "KiB")
        os.remove(tmp)


if __name__ == "__main__":
    main()
