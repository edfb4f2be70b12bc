"""Test infrastructure for the aspect-solution rows (marxasp): the stock program and descriptor dumper of
oracle/_ref, the plain-C restatement (oracle/aspsol_oracle.c), the committed fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np

from tests.fits_table import read_bintable
from tests.level1_lib import run_stock_marx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
HAVE_REF = os.path.exists(os.path.join(REF, "marxasp")) and os.path.exists(os.path.join(REF, "asp_dump"))
COLS = ("time", "ra", "dec", "roll", "q0", "q1", "q2", "q3")

# marx.par arguments of the simulation whose aspect solution is written, marxasp arguments
CASES = {
    "aspsol_default": (["GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "MinEnergy=1.0", "MaxEnergy=2.0"], []),
    "aspsol_roll_pole": (["GratingType=NONE", "DetectorType=HRC-S", "DitherModel=INTERNAL", "MinEnergy=1.0", "MaxEnergy=2.0",
                          "RA_Nom=10.25", "Dec_Nom=88.5", "Roll_Nom=17.0", "SourceRA=10.25", "SourceDEC=88.5", "DitherAmp_RA=40", "DitherAmp_Dec=25", "DitherAmp_Roll=900",
                          "DitherPeriod_RA=707.1", "DitherPeriod_Dec=1087.3", "DitherPeriod_Roll=331.7", "DitherPhase_RA=0.3",
                          "DitherPhase_Dec=1.1", "DitherPhase_Roll=2.2"], ["TimeDel=0.5"]),
    "aspsol_south_no_dither": (["GratingType=NONE", "DetectorType=ACIS-I", "DitherModel=NONE", "MinEnergy=1.0", "MaxEnergy=2.0",
                                "RA_Nom=359.9", "Dec_Nom=-75.0", "Roll_Nom=300.0", "SourceRA=359.9", "SourceDEC=-75.0"], ["TimeDel=2.05"]),
}


def _env():
    return dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"), USER=os.environ.get("USER", "marx"))


def dump_descriptor(marx_dir, asp_args=()):
    """-> (desc[21], num_rows) as the stock marxasp initialisation derives them (oracle/ref/asp_dump.c)"""
    out = subprocess.run([os.path.join(REF, "asp_dump"), "@@" + os.path.join(REF, "par", "marxasp.par"), "MarxDir=" + str(marx_dir)]
                         + list(asp_args), env=_env(), stdout=subprocess.PIPE, text=True, check=True).stdout
    kv = {ln.split()[0]: [float(v) for v in ln.split()[1:]] for ln in out.splitlines() if ln.strip()}
    desc = np.array(kv["time_start"] + kv["delta_time"] + kv["amp"] + kv["period"] + kv["phase"] + kv["nominal_roll"]
                    + kv["pointing"] + kv["ra_hat"] + kv["dec_hat"], dtype=np.float64)
    assert desc.shape == (21,)
    return desc, int(kv["num_rows"][0])


def run_stock_marxasp(marx_dir, fits_path, asp_args=()):
    """the UNMODIFIED marxasp -> columns of its ASPSOL table"""
    p = subprocess.run([os.path.join(REF, "marxasp"), "@@" + os.path.join(REF, "par", "marxasp.par"), "MarxDir=" + str(marx_dir),
                        "OutputFile=" + str(fits_path)] + list(asp_args), env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    t, _ = read_bintable(str(fits_path), "ASPSOL")
    out = {k: np.ascontiguousarray(t[k].astype(np.float64)) for k in ("time", "ra", "dec", "roll")}
    for j in range(4):
        out["q%d" % j] = np.ascontiguousarray(t["q_att"][:, j].astype(np.float64))
    for k in ("dy", "dz", "dtheta"):
        out[k] = np.ascontiguousarray(t[k].astype(np.float32))
    return out


def stock_case(name, tmpdir, n_rays=20000, seed=3):
    """runs the stock marx + marxasp for one case -> (desc, num_rows, reference columns)"""
    args, asp_args = CASES[name]
    out = os.path.join(str(tmpdir), name)
    run_stock_marx(out, args, n_rays=n_rays, seed=seed)
    desc, num = dump_descriptor(out, asp_args)
    ref = run_stock_marxasp(out, os.path.join(str(tmpdir), name + ".fits"), asp_args)
    return desc, num, ref


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z["desc"], int(z["num_rows"]), int(z["first_row"]), {k: z["ref." + k] for k in COLS}


_lib = None


def oracle_rows(desc, first_row, n):
    """oracle/aspsol_oracle.c -> dict of the 8 double columns"""
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _lib.aspsol_oracle_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    desc = np.ascontiguousarray(desc, dtype=np.float64)
    cols = np.zeros((8, n), dtype=np.float64)
    assert 0 == _lib.aspsol_oracle_rows(desc.ctypes.data, int(first_row), int(n), cols.ctypes.data)
    return {k: cols[j] for j, k in enumerate(COLS)}
