"""Test infrastructure for the pile-up model (marxpileup): the stock program and its replay build in oracle/_ref,
the plain-C restatement (oracle/pileup_oracle.c), the committed fixtures."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

from marx_b200 import read_marx_column
from tests.level1_lib import run_stock_marx
from tests.oracle_lib import Oracle, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
HAVE_REF = os.path.exists(os.path.join(REF, "marxpileup_replay")) and os.path.exists(os.path.join(REF, "marx"))
DITHER = ("sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta")
IN_FILES = {"ccd": "detector.dat", "x": "xpixel.dat", "y": "ypixel.dat", "t": "time.dat", "benergy": "b_energy.dat"}
OUT_FILES = dict(IN_FILES, frame="frame.dat", nphotons="nphotons.dat", pha="pha.dat")

# marx.par arguments of the (bright) simulation, marxpileup arguments, calibration pack whose FEF tables the oracle reads
CASES = {
    "pileup_acis_s_bright": (["GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "MinEnergy=0.5", "MaxEnergy=4.0",
                              "SourceFlux=0.05"], ["Alpha=0.5", "FrameTime=3.2"], "c1_acis_s"),
    # (the stock program cannot read a simulation without dither: it opens sky_ra.dat ... unconditionally, marxpileup.c:62,409-438)
    "pileup_acis_s_moderate": (["GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "DitherAmp_RA=4", "DitherAmp_Dec=2",
                                "MinEnergy=1.0", "MaxEnergy=6.0", "SourceFlux=0.004"],
                               ["Alpha=0.9", "FrameTime=1.7", "FrameTransferTime=0.0"], "c1_acis_s"),
    "pileup_acis_i_offaxis": (["GratingType=NONE", "DetectorType=ACIS-I", "DitherModel=INTERNAL", "MinEnergy=0.5", "MaxEnergy=7.0",
                               "SourceFlux=0.02", "SourceRA=250.2134679741175", "SourceDEC=-53.70"], ["Alpha=0.3", "FrameTime=3.2"],
                              "c4_beta_acis_i"),
}


def _env(**extra):
    return dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"), USER=os.environ.get("USER", "marx"), **extra)


def read_dir(d, files):
    out = {}
    for key, f in files.items():
        path = os.path.join(str(d), f)
        if os.path.exists(path):
            out[key] = np.ascontiguousarray(read_marx_column(path)[1])
    for key in DITHER:
        path = os.path.join(str(d), key + ".dat")
        if os.path.exists(path):
            out[key] = np.ascontiguousarray(read_marx_column(path)[1])
    return out


def pileup_params(asp_args):
    """-> (alpha, frame_time) as marxpileup's initialize derives them (marxpileup.c:1076-1084; par defaults 0.5, 3.2, 0.041)"""
    kv = {"Alpha": 0.5, "FrameTime": 3.2, "FrameTransferTime": 0.041}
    for a in asp_args:
        k, v = a.split("=")
        kv[k] = float(v)
    ft = kv["FrameTime"] + (kv["FrameTransferTime"] if kv["FrameTransferTime"] > 0.0 else 0.0)
    return kv["Alpha"], ft


def run_stock_pileup(marx_dir, pu_args, seed=None):
    """seed None: the stock marxpileup (own RNG); else the replay build with the per-frame Philox stream.  -> output columns"""
    exe = "marxpileup" if seed is None else "marxpileup_replay"
    env = _env() if seed is None else _env(PILEUP_SEED=str(seed))
    p = subprocess.run([os.path.join(REF, exe), "@@" + os.path.join(REF, "par", "marxpileup.par"), "MarxOutputDir=" + str(marx_dir),
                        "Verbose=0"] + list(pu_args), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    return read_dir(os.path.join(str(marx_dir), "pileup"), OUT_FILES)


def stock_case(name, tmpdir, n_rays=200000, seed=5, draw_seed=9):
    args, pu_args, pack = CASES[name]
    out = os.path.join(str(tmpdir), name)
    run_stock_marx(out, args, n_rays=n_rays, seed=seed)
    return read_dir(out, IN_FILES), run_stock_pileup(out, pu_args, seed=draw_seed), out


_ready = False


def oracle_pileup(cols, pu_args, pack, seed):
    """oracle/pileup_oracle.c on the input columns -> output columns"""
    global _ready
    L = lib()
    if not _ready:
        L.pileup_oracle_run.restype = C.c_longlong
        L.pileup_oracle_run.argtypes = ([C.c_void_p, C.c_uint64] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_uint64, C.c_uint64]
                                        + [C.c_void_p] * 9)
        _ready = True
    alpha, frame_time = pileup_params(pu_args)
    n = len(cols["t"])
    have_d = all(k in cols for k in DITHER)
    ccd = np.ascontiguousarray(cols["ccd"], dtype=np.int8)
    f = {k: np.ascontiguousarray(cols[k], dtype=np.float32) for k in ("x", "y", "t", "benergy")}
    keep = [np.ascontiguousarray(cols[k], dtype=np.float32) for k in DITHER] if have_d else []
    din = (C.c_void_p * 6)(*[a.ctypes.data for a in keep]) if have_d else None
    out = {"ccd": np.zeros(n, np.int8), "x": np.zeros(n, np.float32), "y": np.zeros(n, np.float32), "t": np.zeros(n, np.float32),
           "benergy": np.zeros(n, np.float32), "frame": np.zeros(n, np.int32), "nphotons": np.zeros(n, np.int16), "pha": np.zeros(n, np.int16)}
    dout_arrays = [np.zeros(n, np.float32) for _ in DITHER] if have_d else []
    dout = (C.c_void_p * 6)(*[a.ctypes.data for a in dout_arrays]) if have_d else None
    o = Oracle(pack, 0)
    m = L.pileup_oracle_run(o._h, n, ccd.ctypes.data, f["x"].ctypes.data, f["y"].ctypes.data, f["t"].ctypes.data, f["benergy"].ctypes.data,
                            din, alpha, frame_time, int(seed), n,
                            out["ccd"].ctypes.data, out["x"].ctypes.data, out["y"].ctypes.data, out["t"].ctypes.data,
                            out["benergy"].ctypes.data, out["frame"].ctypes.data, out["nphotons"].ctypes.data, out["pha"].ctypes.data, dout)
    o.close()
    assert m >= 0, "pileup_oracle_run failed"
    res = {k: v[:m] for k, v in out.items()}
    for k, a in zip(DITHER, dout_arrays):
        res[k] = a[:m]
    return res


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cols = {k[3:]: z[k] for k in z.files if k.startswith("in.")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref.")}
    return cols, ref, int(z["draw_seed"])
