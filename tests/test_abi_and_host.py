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
    assert marx_b200.load_library().marxb200_abi_version() == 2


@pytest.mark.skipif(not _built(), reason="libmarxb200.so not built")
def test_shard_of_partitions_the_ray_range():
    """marxb200_shard_of (the block rule of marxb200_trace_sharded): contiguous, super-tile aligned, ragged/empty at the end"""
    from marx_b200.api import shard_of
    for first, n_total, world in [(0, 1 << 24, 8), (12345 * 65536, 10**7, 8), (7, 1000, 2), (0, 65536 * 3 + 5, 4), (1 << 40, 10**9, 3),
                                  (0, 1, 8), (0, 65536 * 8, 8), (0, 65536 * 8 + 1, 8)]:
        blocks = [shard_of(first, n_total, r, world) for r in range(world)]
        pos = first
        for f, n in blocks:
            assert f >= pos and (f == pos or n == 0)          # contiguous; empty blocks sit at the end
            assert (f - first) % 65536 == 0                    # blocks start on super-tile boundaries of the step
            pos = f + n if n else pos
        assert sum(n for _, n in blocks) == n_total and pos == first + n_total
        sizes = [n for _, n in blocks if n]
        assert all(s == sizes[0] for s in sizes[:-1]) and sizes[-1] <= sizes[0] and sizes[0] % 65536 == 0 or len(sizes) == 1
    with pytest.raises(marx_b200.MarxB200Error):
        shard_of(0, 10, 3, 2)


@pytest.mark.skipif(not _built(), reason="libmarxb200.so not built")
def test_comm_unique_id_needs_no_gpu():
    """the id hand-over half of marxb200_comm_init works on a CPU box (NCCL is opened with dlopen at run time)"""
    from marx_b200.api import comm_unique_id, COMM_ID_BYTES
    try:
        a, b = comm_unique_id(), comm_unique_id()
    except marx_b200.MarxB200Error as e:
        pytest.skip("no NCCL library on this box: %s" % e)
    assert len(a) == len(b) == COMM_ID_BYTES and a != b


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
