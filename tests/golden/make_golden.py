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
}


def main():
    par = "@@" + os.path.join(REF, "par", "marx.par")
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    for name, cfg in CONFIGS.items():
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
