"""CPU-side checks: the C-ABI library loads and exports every symbol include/marxb200.h declares, fails
loudly without a GPU, the calibration packs parse, and the photon record layout matches the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import marx_b200
from marx_b200.api import EXPORTED_SYMBOLS, PHOTON_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    return os.path.exists(marx_b200.lib_path())


def test_header_symbols_are_listed():
    hdr = open(os.path.join(ROOT, "include", "marxb200.h")).read()
    declared = set(re.findall(r"\b(marxb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(EXPORTED_SYMBOLS), declared ^ set(EXPORTED_SYMBOLS)


@pytest.mark.skipif(not _built(), reason="libmarxb200.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = C.CDLL(marx_b200.lib_path())
    for name in EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name
    assert marx_b200.load_library().marxb200_abi_version() == 1


@pytest.mark.skipif(not _built(), reason="libmarxb200.so not built")
def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(marx_b200.MarxB200Error, match="no CPU fallback"):
        marx_b200.MarxB200("c2_hetg_acis_s")


def test_photon_record_layout_matches_reference():
    # sizeof(Marx_Photon_Attr_Type) == 136 with these offsets (probed on the reference build)
    assert PHOTON_DTYPE.itemsize == 136
    off = {n: PHOTON_DTYPE.fields[n][1] for n in PHOTON_DTYPE.names}
    assert (off["energy"], off["x"], off["p"], off["arrival_time"], off["flags"]) == (0, 8, 32, 56, 64)
    assert (off["y_pixel"], off["dither"], off["pi"], off["pulse_height"], off["mirror_shell"]) == (68, 84, 108, 112, 116)
    assert (off["ccd_num"], off["detector_region"], off["order"], off["support_orders"], off["tag"]) == (120, 121, 122, 123, 128)


@pytest.mark.parametrize("name", ["c1_acis_s", "c2_hetg_acis_s"])
def test_calpack_contents(name):
    import struct
    b = open(marx_b200.caldata_path(name), "rb").read()
    assert b[:8] == b"MXB2CAL1"
    n, = struct.unpack_from("<I", b, 8)
    off, names = 16, {}
    for _ in range(n):
        nm = b[off:off + 56].split(b"\0")[0].decode()
        dtype, _, count = struct.unpack_from("<IIQ", b, off + 56)
        off += 72
        names[nm] = (dtype, count, off)
        off += (count * (8 if dtype == 0 else 4) + 7) // 8 * 8
    assert off == len(b)
    for need in ["source.params", "dither.params", "hrma.params", "hrma.opt_energies", "grating.params", "acis.params",
                 "hrma.shell3.wfold_h.theta", "acis.chip5.fef_map", "acis.num_fefs"]:
        assert need in names, need
    # cumulative aperture fractions end at exactly 1 (hrma.c:739-744)
    d, c, o = names["hrma.shell3.params"]
    assert np.frombuffer(b, "<f8", c, o)[19] == 1.0
    if name.startswith("c2"):
        d, c, o = names["grating.shell0.cum_eff"]
        ce = np.frombuffer(b, "<f4", c, o).reshape(23, -1)
        assert (np.diff(ce, axis=0) >= 0).all() and ce.max() <= 1.0 + 1e-6     # diffract.c:1284-1297
