"""Level-1 test infrastructure: drive the stock marx2fits of oracle/_ref (replay RNG), parse its descriptor dump and its
EVENTS table, call the plain-C oracle (oracle/level1_oracle.c) through ctypes.  Import from tests only."""
import ctypes as C
import os
import subprocess

import numpy as np

from marx_b200.api import read_marx_column
from marx_b200.level1 import DETECTOR_TYPES, LEVEL1_COLUMNS, Level1Desc, _Desc, _Level1Columns, alloc_columns
from tests import oracle_lib
from tests.fits_table import read_bintable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (marx.par overrides, marx2fits --pixadj, draws per event row)
LEVEL1_CASES = {
    "level1_acis_s_hetg_edser": (["SourceType=POINT", "MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S",
                                  "DitherModel=INTERNAL"], "edser", 1),
    "level1_acis_i_beta_randomize": (["SourceType=BETA", "S-BetaCoreRadius=10", "S-BetaBeta=0.7", "SourceDEC=-53.92410480125", "MinEnergy=0.5",
                                      "MaxEnergy=7.0", "GratingType=NONE", "DetectorType=ACIS-I", "DitherModel=INTERNAL"], "randomize", 3),
    "level1_acis_s_nodither_none": (["SourceType=POINT", "MinEnergy=1.5", "MaxEnergy=1.5", "GratingType=NONE", "DetectorType=ACIS-S",
                                     "DitherModel=NONE"], "none", 1),
    "level1_acis_s_hetg_exact": (["SourceType=POINT", "MinEnergy=0.3", "MaxEnergy=8.0", "GratingType=HETG", "DetectorType=ACIS-S",
                                  "DitherModel=INTERNAL", "DetOffsetX=0.3", "DetOffsetZ=-1.5"], "exact", 1),
    "level1_hrc_s_letg": (["SourceType=POINT", "MinEnergy=0.1", "MaxEnergy=2.0", "GratingType=LETG", "DetectorType=HRC-S",
                           "DitherModel=INTERNAL"], "edser", 2),
}
# HRC-I: the stock marx2fits of this checkout stops with "DetectorType HRC-I not supported" (marx_get_detector_info fails in its
# own initialisation), so there is no reference output to pin that detector's Level-1 columns against.
COMMON = ["ExposureTime=0", "Verbose=0", "SourceFlux=0.003", "TStart=2023.5", "SpectrumType=FLAT"]
INPUT_FILES = {"time": "time.dat", "xpixel": "xpixel.dat", "ypixel": "ypixel.dat", "b_energy": "b_energy.dat", "hrc_u": "hrc_u.dat",
               "hrc_v": "hrc_v.dat", "pha": "pha.dat", "ccd": "detector.dat", "sky_ra": "sky_ra.dat", "sky_dec": "sky_dec.dat",
               "sky_roll": "sky_roll.dat", "det_dy": "det_dy.dat", "det_dz": "det_dz.dat", "det_theta": "det_theta.dat"}
DITHER_KEYS = ["sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta"]
# EVENTS column of the stock marx2fits -> marxb200_level1_columns member
FITS_TO_L1 = {"TIME": "time", "CCD_ID": "ccd_id", "CHIP_ID": "ccd_id", "NODE_ID": "node_id", "EXPNO": "expno", "CHIPX": "chipx", "CHIPY": "chipy",
              "TDETX": "tdetx", "TDETY": "tdety", "DETX": "detx", "DETY": "dety", "X": "x", "Y": "y", "PHA": "pha", "ENERGY": "energy",
              "PI": "pi", "FLTGRADE": "fltgrade", "GRADE": "grade", "STATUS": "status", "U": "hrc_u", "V": "hrc_v"}


def have_reference():
    return all(os.path.exists(os.path.join(REF, b)) for b in ("marx", "marx2fits_replay", "level1_dump"))


def _env(**extra):
    return dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"), **extra)


def run_stock_marx(outdir, args, n_rays=200000, seed=7):
    cmd = [os.path.join(REF, "marx"), "@@" + os.path.join(REF, "par", "marx.par"), "OutputDir=" + str(outdir),
           "OutputVectors=#ETXYZ123DxyMPOabcdSrB"] + COMMON + list(args) + ["NumRays=%d" % n_rays, "dNumRays=100000", "RandomSeed=%d" % seed]
    p = subprocess.run(cmd, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:]


def read_inputs(outdir):
    """the column files marx2fits reads, as the float32/int16/int8 values they hold"""
    cols = {}
    for key, f in INPUT_FILES.items():
        path = os.path.join(str(outdir), f)
        if os.path.exists(path):
            cols[key] = np.ascontiguousarray(read_marx_column(path)[1])
    return cols


def read_subpix(path):
    """acis_subpix.c:106-197 (read_subpix_ext) + :217-243: -> npoints[2*256], offset[2*256], data"""
    npoints = np.zeros(2 * 256, dtype=np.int32)
    offset = np.zeros(2 * 256, dtype=np.uint32)
    data = []
    pos = 0
    for t, ext in enumerate(("MARX_ACIS_SUBPIX_FI", "MARX_ACIS_SUBPIX_BI")):
        tab, _ = read_bintable(path, ext)
        tab = {k.upper(): v for k, v in tab.items()}
        for r in range(len(tab["FLTGRADE"])):
            g, n = int(tab["FLTGRADE"][r]), int(tab["NPOINTS"][r])
            if n <= 0:
                continue
            npoints[t * 256 + g], offset[t * 256 + g] = n, pos
            for col in ("ENERGY", "CHIPX_OFFSET", "CHIPY_OFFSET"):
                data.append(np.asarray(tab[col][r][:n], dtype=np.float32))
            pos += 3 * n
    return npoints, offset, (np.concatenate(data) if data else np.zeros(0, dtype=np.float32))


def dump_descriptor(outdir, pixadj):
    """the values the stock marx2fits initialisation derives for this output directory (oracle/_ref/level1_dump)"""
    p = subprocess.run([os.path.join(REF, "level1_dump"), "--pixadj=" + pixadj, str(outdir)], env=_env(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    d, chips = {}, []
    for line in p.stdout.splitlines():
        key, _, rest = line.partition(" ")
        vals = rest.split()
        if key == "chip":
            chips.append([float(v) for v in vals])
        elif key == "detector":
            d["detector_type"] = DETECTOR_TYPES[vals[0]]
        elif key in ("fp", "det_offset"):
            d[key] = np.array([float(v) for v in vals])
        elif key == "subpix_file":
            d["subpix_file"] = vals[0]
        elif key in ("used_dither", "pix_adjust"):
            d[key] = int(vals[0])
        elif key in ("time_del", "time_start", "pi_factor", "focal_length", "nominal_roll"):
            d[key] = float(vals[0])
    d["chips"] = np.array(chips)
    if "subpix_file" in d and d["pix_adjust"] == 2:
        f = d.pop("subpix_file")
        f = f if os.path.isabs(f) else os.path.join(REF, "data", f)
        d["subpix_npoints"], d["subpix_offset"], d["subpix_data"] = read_subpix(f)
    d.pop("subpix_file", None)
    return d


def run_stock_marx2fits(outdir, fits_path, pixadj, ndraw, seed):
    """the UNMODIFIED marx2fits.c with the per-row Philox stream (oracle/ref/level1_rng.c) -> EVENTS columns"""
    p = subprocess.run([os.path.join(REF, "marx2fits_replay"), "--pixadj=" + pixadj, str(outdir), str(fits_path)],
                       env=_env(L1_NDRAW=str(ndraw), L1_SEED=str(seed)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    ev, _ = read_bintable(str(fits_path), "EVENTS")
    out = {}
    for name, arr in ev.items():
        if name == "STATUS":                       # 32X: 4 bytes, big endian bit string
            arr = (arr.astype(np.uint32) << np.array([24, 16, 8, 0], dtype=np.uint32)).sum(axis=1).astype(np.int64)
        out[name] = arr
    return out


# ---- the plain-C oracle ----
class _State(C.Structure):
    _fields_ = [("last_expno", C.c_int64), ("update_dither", C.c_int32), ("dither", C.c_float * 6), ("rows", C.c_uint64)]


class _Input(C.Structure):
    _fields_ = [("n", C.c_uint64)] + [(k, C.c_void_p) for k in ("time", "xpixel", "ypixel", "b_energy", "hrc_u", "hrc_v")] + \
               [("dither", C.c_void_p * 6), ("pha", C.c_void_p), ("ccd", C.c_void_p)]


class Level1Oracle:
    def __init__(self, desc, seed):
        self.lib = oracle_lib.lib()
        self.lib.level1_oracle_transform.restype = C.c_long
        self.lib.level1_oracle_transform.argtypes = [C.POINTER(_Desc), C.c_uint64, C.POINTER(_State), C.POINTER(_Input), C.POINTER(_Level1Columns)]
        self.lib.level1_oracle_reset.argtypes = [C.POINTER(_State)]
        self.desc = desc if isinstance(desc, Level1Desc) else Level1Desc.from_dict(desc)
        if self.desc.c.detector_type < 3 and self.desc.c.pix_adjust == 2:
            self.desc.c.pix_adjust = 1             # marx2fits main :3308-3309 (the C ABI does the same in marxb200_set_level1)
        self.seed = int(seed)
        self.state = _State()
        self.lib.level1_oracle_reset(C.byref(self.state))

    def transform(self, cols):
        """cols: dict of input columns (read_inputs / GPU egress columns) -> dict of Level-1 columns"""
        n = len(cols["time"])
        inp = _Input()
        inp.n = n
        keep = []

        def ptr(key, dt):
            if key not in cols or cols[key] is None:
                return None
            a = np.ascontiguousarray(cols[key], dtype=dt)
            keep.append(a)
            return a.ctypes.data
        for k in ("time", "xpixel", "ypixel", "b_energy", "hrc_u", "hrc_v"):
            setattr(inp, k, ptr(k, np.float32))
        for j, k in enumerate(DITHER_KEYS):
            inp.dither[j] = ptr(k, np.float32)
        inp.pha = ptr("pha", np.int16)
        inp.ccd = ptr("ccd", np.int8)
        out_c, out = alloc_columns(n)
        got = self.lib.level1_oracle_transform(C.byref(self.desc.c), self.seed, C.byref(self.state), C.byref(inp), C.byref(out_c))
        if got != n:
            raise RuntimeError("level1 oracle rejected row %d" % (-got - 1))
        return {k: v[:n] for k, v in out.items()}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    desc = {k[5:]: z[k] for k in z.files if k.startswith("desc.")}
    cols = {k[3:]: z[k] for k in z.files if k.startswith("in.")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref.")}
    return desc, cols, ref, int(z["seed"])


def photons_from_columns(cols):
    """the device list a marx run would hold for these column files: records whose file narrowing reproduces them exactly"""
    from marx_b200.api import PHOTON_DTYPE
    n = len(cols["time"])
    ph = np.zeros(n, dtype=PHOTON_DTYPE)
    ph["arrival_time"] = cols["time"].astype(np.float64)
    ph["y_pixel"], ph["z_pixel"] = cols["xpixel"], cols["ypixel"]
    if "b_energy" in cols:
        ph["pi"] = cols["b_energy"]
    if "hrc_u" in cols:
        ph["u_pixel"], ph["v_pixel"] = cols["hrc_u"], cols["hrc_v"]
    ph["pulse_height"] = cols["pha"]
    ph["ccd_num"] = cols["ccd"]
    for j, k in enumerate(DITHER_KEYS):
        if k in cols:
            ph["dither"][:, j] = cols[k]
    ph["tag"] = np.arange(n, dtype=np.uint32)
    return ph


def sky_y_tolerance(y_ref, desc, ulps=4.0):
    """How far a correct Y may be from the reference's: marx_mnc_to_ra_dec (pixlib.c:503-527) takes dec = acos (perp) with
    perp = sqrt (x^2 + y^2) of a unit vector within ~1e-3 rad of the axis, i.e. within 1e-6 of 1.  acos is ill-conditioned there:
    d(perp) = dec * d(dec) + d(dec)^2 / 2, so a perturbation of `ulps` units in the last place of perp (2^-53 each, what a
    1-ulp difference of any sin/cos upstream produces) moves dec by -dec + sqrt (dec^2 + 2E), E = ulps * 2^-53: 6e-3 sky pixels on
    the axis, 8e-5 / |Y - Y0| pixels away from it (measured on B200: |dY| * |Y - Y0| <= 3.93e-5 = 2 such units; 4 allowed).
    Returns pixels."""
    ds0 = float(np.asarray(desc["fp"])[0])
    r = np.abs(np.asarray(y_ref, dtype=np.float64) - float(np.asarray(desc["fp"])[2]))
    e = 2.0 * ulps * 2.0 ** -53 / (ds0 * ds0)
    return -r + np.sqrt(r * r + e) + 1e-9


def compare_with_fits(l1, fits, f32_ulps=0, y_tolerance=None):
    """Level-1 columns (oracle or GPU; every row) against the EVENTS table of the stock marx2fits (kept rows only).
    f32_ulps = 0: bit for bit after the writer's casts (the oracle's bar); > 0: float32 columns may differ by that many ulps
    (GPU libm), integers stay exact."""
    keep = l1["keep"].astype(bool)
    report = {}
    for fname, arr in fits.items():
        if fname not in FITS_TO_L1:
            continue
        mine = l1[FITS_TO_L1[fname]][keep]
        assert len(mine) == len(arr), (fname, len(mine), len(arr))
        if fname == "Y" and y_tolerance is not None:
            # float32 rounding of both sides (ulp 4.9e-4 at 4096) on top of the conditioning bound
            tol = y_tolerance[keep] + 2.0 * np.spacing(np.abs(arr).astype(np.float32)).astype(np.float64)
            assert (np.abs(mine.astype(np.float32).astype(np.float64) - arr) <= tol).all(), fname
            report[fname] = 0
        elif arr.dtype.kind == "f" and arr.dtype.itemsize == 4:
            a, b = mine.astype(np.float32), arr.astype(np.float32)       # write_float64_as_float32 / write_float32
            d = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
            report[fname] = int(d.max()) if len(d) else 0
            assert report[fname] <= f32_ulps, (fname, report[fname])
        elif arr.dtype.kind == "f":
            if f32_ulps == 0:
                assert np.array_equal(mine.astype(np.float64), arr), fname
            else:
                assert np.allclose(mine, arr, rtol=1e-12, atol=0), fname
            report[fname] = 0
        else:
            assert np.array_equal(mine.astype(np.int64), arr.astype(np.int64)), (fname, np.flatnonzero(mine.astype(np.int64) != arr.astype(np.int64))[:5])
            report[fname] = 0
    return report
