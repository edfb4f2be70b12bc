"""ctypes mirror of the Level-1 part of include/marxb200.h (marxb200_level1_*): the per-event transforms of marx2fits
(marx/src/marx2fits.c:3584-3943) on the device-resident event list.  `Level1Desc.from_dict` takes the values the stock
marx2fits initialisation leaves in its statics (what oracle/ref/level1_dump.c prints, what tests/golden/level1_*.npz
hold); in a MARX build they are filled in by marx2fits itself."""
import ctypes as C

import numpy as np

STAGE_LEVEL1 = 4
MAX_CHIPS = 10
PIXADJ = {"none": 0, "randomize": 1, "edser": 2, "exact": 3}
DETECTOR_TYPES = {"HRC-S": 1, "HRC-I": 2, "ACIS-S": 3, "ACIS-I": 4}


class _Chip(C.Structure):
    _fields_ = [("id", C.c_int32), ("subpix_table", C.c_int32), ("x_ll", C.c_double * 3), ("xhat", C.c_double * 3),
                ("yhat", C.c_double * 3), ("x_pixel_size", C.c_double), ("y_pixel_size", C.c_double),
                ("xpixel_offset", C.c_double), ("ypixel_offset", C.c_double), ("tdet_xoff", C.c_float), ("tdet_yoff", C.c_float)]


class _Desc(C.Structure):
    _fields_ = [("detector_type", C.c_int32), ("num_chips", C.c_int32), ("chips", _Chip * MAX_CHIPS),
                ("fp_delta_s0", C.c_double), ("fp_x0", C.c_double), ("fp_y0", C.c_double), ("focal_length", C.c_double),
                ("det_offset", C.c_double * 3), ("time_del", C.c_double), ("time_start", C.c_double), ("pi_factor", C.c_double),
                ("nominal_roll", C.c_double), ("used_dither", C.c_int32), ("pix_adjust", C.c_int32),
                ("subpix_npoints", C.c_void_p), ("subpix_offset", C.c_void_p), ("subpix_data", C.c_void_p),
                ("subpix_data_len", C.c_uint64)]


# marxb200_level1_columns, in declaration order
LEVEL1_COLUMNS = [("time", "<f8"), ("detx", "<f8"), ("dety", "<f8"), ("x", "<f8"), ("y", "<f8"),
                  ("expno", "<i4"), ("tdetx", "<i4"), ("tdety", "<i4"), ("pha", "<i4"), ("hrc_u", "<i4"), ("hrc_v", "<i4"),
                  ("energy", "<f4"),
                  ("ccd_id", "<i2"), ("node_id", "<i2"), ("chipx", "<i2"), ("chipy", "<i2"), ("pi", "<i2"), ("fltgrade", "<i2"),
                  ("grade", "<i2"), ("status", "<i2"), ("keep", "u1")]


class _Level1Columns(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in LEVEL1_COLUMNS]


def alloc_columns(n, names=None):
    """-> (ctypes struct, dict of numpy arrays) for n rows"""
    cols, arrays = _Level1Columns(), {}
    for name, dt in LEVEL1_COLUMNS:
        if names is not None and name not in names:
            continue
        arrays[name] = np.zeros(max(int(n), 1), dtype=dt)
        setattr(cols, name, arrays[name].ctypes.data)
    return cols, arrays


class Level1Desc:
    """Owns the numpy arrays the C descriptor points into."""

    def __init__(self):
        self.c = _Desc()
        self._keep = []

    @classmethod
    def from_dict(cls, d):
        """d: detector_type, chips (n x 16: id, x_ll[3], xhat[3], yhat[3], x/y_pixel_size, x/ypixel_offset, tdet_x/yoff), fp (3),
        focal_length, det_offset (3), time_del, time_start, pi_factor, nominal_roll, used_dither, pix_adjust and, for EDSER,
        subpix_npoints (512), subpix_offset (512), subpix_data."""
        self = cls()
        c = self.c
        c.detector_type = int(d["detector_type"])
        chips = np.asarray(d["chips"], dtype=np.float64).reshape(-1, 16)
        c.num_chips = len(chips)
        for k, row in enumerate(chips):
            g = c.chips[k]
            g.id = int(row[0])
            g.subpix_table = 1 if int(row[0]) in (5, 7) else 0        # acis_subpix.c:262-268
            for j in range(3):
                g.x_ll[j], g.xhat[j], g.yhat[j] = row[1 + j], row[4 + j], row[7 + j]
            g.x_pixel_size, g.y_pixel_size, g.xpixel_offset, g.ypixel_offset = row[10], row[11], row[12], row[13]
            g.tdet_xoff, g.tdet_yoff = row[14], row[15]
        fp = np.asarray(d["fp"], dtype=np.float64)
        c.fp_delta_s0, c.fp_x0, c.fp_y0 = fp[0], fp[1], fp[2]
        c.focal_length = float(d["focal_length"])
        for j in range(3):
            c.det_offset[j] = float(np.asarray(d["det_offset"])[j])
        c.time_del, c.time_start = float(d["time_del"]), float(d["time_start"])
        c.pi_factor, c.nominal_roll = float(d["pi_factor"]), float(d["nominal_roll"])
        c.used_dither, c.pix_adjust = int(d["used_dither"]), int(d["pix_adjust"])
        if "subpix_data" in d and len(np.asarray(d["subpix_data"])):
            npts = np.ascontiguousarray(d["subpix_npoints"], dtype=np.int32)
            offs = np.ascontiguousarray(d["subpix_offset"], dtype=np.uint32)
            data = np.ascontiguousarray(d["subpix_data"], dtype=np.float32)
            self._keep += [npts, offs, data]
            c.subpix_npoints, c.subpix_offset, c.subpix_data = npts.ctypes.data, offs.ctypes.data, data.ctypes.data
            c.subpix_data_len = len(data)
        return self

    def byref(self):
        return C.byref(self.c)
