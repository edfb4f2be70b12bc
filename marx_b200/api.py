"""ctypes host mirror of include/marxb200.h.

Method names follow the reference's module API (marx/libsrc/marx.h:287-293,354-362):
``create_photons`` / ``mirror_reflect`` / ``grating_diffract`` / ``detect`` over a photon list that
lives in HBM; ``download`` returns the live list as records laid out exactly like the reference's
``Marx_Photon_Attr_Type`` (marx.h:51-100).  Error behaviour mirrors the reference: the C calls return
-1 and a message; here that raises ``MarxB200Error``.
"""
import ctypes as C
import os

import numpy as np

STAGE_SOURCE, STAGE_MIRROR, STAGE_GRATING, STAGE_DETECTOR = 0, 1, 2, 3

# every symbol include/marxb200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "marxb200_abi_version", "marxb200_last_error", "marxb200_create", "marxb200_destroy", "marxb200_set_stream",
    "marxb200_set_compaction", "marxb200_set_source", "marxb200_set_dither", "marxb200_set_hrma", "marxb200_set_flatfield",
    "marxb200_set_grating", "marxb200_set_acis", "marxb200_set_hrc_s", "marxb200_load_calpack", "marxb200_alloc_photons",
    "marxb200_create_photons", "marxb200_truncate_exposure", "marxb200_time_sums", "marxb200_mirror_reflect", "marxb200_grating_diffract",
    "marxb200_detect", "marxb200_restore_order", "marxb200_trace", "marxb200_trace_from", "marxb200_set_profiling", "marxb200_get_kernel_ms", "marxb200_get_counts", "marxb200_get_stage_counts", "marxb200_get_internal_counts", "marxb200_download",
    "marxb200_upload", "marxb200_upload_from", "marxb200_download_all", "marxb200_download_columns", "marxb200_egress_begin", "marxb200_egress_end", "marxb200_write_photons", "marxb200_set_async_writer", "marxb200_write_flush", "marxb200_egress_begin_packed", "marxb200_egress_end_packed", "marxb200_measure_fp64_peak", "marxb200_get_launch_count",
    "marxb200_tally_create", "marxb200_tally_accumulate", "marxb200_tally_reset", "marxb200_tally_read", "marxb200_tally_device_ptr",
    "marxb200_set_level1", "marxb200_level1_reset", "marxb200_level1_transform", "marxb200_level1_download",
    "marxb200_aspsol_rows",
    "marxb200_pileup_run", "marxb200_pileup_events",
    "marxb200_comm_get_unique_id", "marxb200_comm_init", "marxb200_comm_init_file", "marxb200_comm_info", "marxb200_comm_destroy",
    "marxb200_shard_of", "marxb200_trace_sharded", "marxb200_tally_allreduce",
    "marxb200_merge_events_begin", "marxb200_merge_events_end", "marxb200_merge_download", "marxb200_probe_d2h", "marxb200_device_warmup",
]

# marxb200_tally_axis.column (include/marxb200.h)
TALLY_COLUMNS = {"energy": 0, "time": 1, "pha": 2, "pi": 3, "order": 4, "ccd": 5, "shell": 6, "chipx": 7, "chipy": 8,
                 "ypos": 9, "zpos": 10}

# Marx_Photon_Attr_Type, marx/libsrc/marx.h:51-100 (136 bytes; offsets probed in SURVEY.md 8a1)
PHOTON_DTYPE = np.dtype({
    "names": ["energy", "x", "p", "arrival_time", "flags", "y_pixel", "z_pixel", "u_pixel", "v_pixel",
              "dither", "pi", "pulse_height", "mirror_shell", "ccd_num", "detector_region", "order",
              "support_orders", "tag"],
    "formats": ["<f8", ("<f8", 3), ("<f8", 3), "<f8", "<u4", "<f4", "<f4", "<f4", "<f4",
                ("<f4", 6), "<f4", "<i2", "<u4", "i1", "i1", "i1", ("i1", 4), "<u4"],
    "offsets": [0, 8, 32, 56, 64, 68, 72, 76, 80, 84, 108, 112, 116, 120, 121, 122, 123, 128],
    "itemsize": 136,
})


# history / write-mask bits, marx/libsrc/marx.h:126-147
HISTORY = {"ENERGY": 0x1, "TIME": 0x2, "X_VECTOR": 0x4, "P_VECTOR": 0x8, "TAG": 0x10, "PULSEHEIGHT": 0x20, "PI": 0x40,
           "DET_PIXEL": 0x80, "DET_NUM": 0x100, "DET_REGION": 0x200, "DET_UV_PIXEL": 0x400, "MIRROR_SHELL": 0x800,
           "SKY_DITHER": 0x1000, "DET_DITHER": 0x2000, "ORDER": 0x100000, "ORDER1": 0x200000, "ORDER2": 0x400000,
           "ORDER3": 0x800000, "ORDER4": 0x1000000}


def read_marx_column(path):
    """Read one column file of a MARX output directory (format: marxio.c:151-205 header + big-endian data).
    Returns (column name, numpy array in native byte order)."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 32 or raw[:4] != bytes([0x83, 0x13, 0x89, 0x8D]):
        raise MarxB200Error("%s: not a MARX column file" % path)
    kind = chr(raw[4])
    name = raw[5:20].split(b"\0")[0].decode()
    rows = int.from_bytes(raw[20:24], "big", signed=True)
    dt = {"E": ">f4", "D": ">f8", "J": ">i4", "I": ">i2", "A": "i1"}[kind]
    data = np.frombuffer(raw, dtype=dt, offset=32)
    if len(data) != rows:
        raise MarxB200Error("%s: header says %d rows, file holds %d" % (path, rows, len(data)))
    return name, data.astype(data.dtype.newbyteorder("="))


class MarxB200Error(RuntimeError):
    pass


def lib_path():
    if os.environ.get("MARXB200_LIB"):          # developer A/B builds; the default is the in-tree library
        return os.environ["MARXB200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmarxb200.so")


def caldata_path(name):
    """Path of a calibration pack shipped with the package (marx_b200/caldata/<name>.calpack)."""
    if not name.endswith(".calpack"):
        name += ".calpack"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "caldata", name)


_LIB = None


def load_library():
    """dlopen libmarxb200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MarxB200Error("libmarxb200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                            "There is no CPU fallback." % path)
    lib = C.CDLL(path)
    u64, i32, vp, dbl = C.c_uint64, C.c_int, C.c_void_p, C.c_double
    lib.marxb200_abi_version.restype = i32
    lib.marxb200_last_error.restype = C.c_char_p
    sigs = {
        "marxb200_create": [C.POINTER(vp), i32, u64],
        "marxb200_destroy": [vp],
        "marxb200_set_stream": [vp, vp],
        "marxb200_set_compaction": [vp, i32],
        "marxb200_load_calpack": [vp, C.c_char_p],
        "marxb200_alloc_photons": [vp, u64],
        "marxb200_create_photons": [vp, u64, u64, dbl],
        "marxb200_truncate_exposure": [vp, dbl, C.POINTER(u64)],
        "marxb200_time_sums": [vp, u64, u64, C.POINTER(dbl), u64, C.POINTER(u64)],
        "marxb200_mirror_reflect": [vp],
        "marxb200_grating_diffract": [vp],
        "marxb200_detect": [vp],
        "marxb200_restore_order": [vp],
        "marxb200_trace": [vp, u64, u64],
        "marxb200_trace_from": [vp, u64, u64, dbl],
        "marxb200_set_profiling": [vp, i32],
        "marxb200_get_kernel_ms": [vp, C.POINTER(dbl), C.POINTER(u64)],
        "marxb200_get_counts": [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(dbl)],
        "marxb200_get_stage_counts": [vp, C.POINTER(u64)],
        "marxb200_get_internal_counts": [vp, C.POINTER(u64)],
        "marxb200_download": [vp, vp, u64, C.POINTER(u64)],
        "marxb200_download_all": [vp, vp, u64, C.POINTER(u64)],
        "marxb200_upload": [vp, vp, u64, vp],
        "marxb200_upload_from": [vp, vp, u64, vp, dbl],
        "marxb200_download_columns": [vp, vp, u64, C.POINTER(u64)],
        "marxb200_get_launch_count": [vp, C.POINTER(u64)],
        "marxb200_aspsol_rows": [vp, vp, u64, u64, vp, vp, C.POINTER(dbl)],
        "marxb200_pileup_run": [vp, u64, vp, dbl, dbl, u64, u64, vp, C.POINTER(u64), C.POINTER(dbl)],
        "marxb200_pileup_events": [vp, dbl, dbl, dbl, u64, u64, vp, C.POINTER(u64), C.POINTER(dbl)],
        "marxb200_egress_begin": [vp, u64],
        "marxb200_egress_end": [vp, vp, C.POINTER(u64)],
        "marxb200_write_photons": [vp, C.c_char_p, u64, i32, dbl],
        "marxb200_set_async_writer": [vp, i32],
        "marxb200_write_flush": [vp],
        "marxb200_measure_fp64_peak": [vp, C.POINTER(dbl)],
        "marxb200_egress_begin_packed": [vp, u64, dbl, u64],
        "marxb200_egress_end_packed": [vp, vp, u64, vp],
        "marxb200_tally_create": [vp, vp, i32],
        "marxb200_tally_accumulate": [vp, i32],
        "marxb200_tally_reset": [vp, i32],
        "marxb200_tally_read": [vp, i32, vp, u64],
        "marxb200_tally_device_ptr": [vp, i32, C.POINTER(vp), C.POINTER(u64)],
        "marxb200_set_level1": [vp, vp],
        "marxb200_level1_reset": [vp],
        "marxb200_level1_transform": [vp, dbl],
        "marxb200_level1_download": [vp, vp, u64, C.POINTER(u64)],
        "marxb200_comm_get_unique_id": [vp],
        "marxb200_comm_init": [vp, vp, i32, i32],
        "marxb200_comm_init_file": [vp, C.c_char_p, i32, i32, dbl],
        "marxb200_comm_info": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        "marxb200_comm_destroy": [vp],
        "marxb200_shard_of": [u64, u64, i32, i32, C.POINTER(u64), C.POINTER(u64)],
        "marxb200_trace_sharded": [vp, u64, u64, dbl, C.POINTER(u64), C.POINTER(u64)],
        "marxb200_tally_allreduce": [vp, i32],
        "marxb200_merge_events_begin": [vp, u64, dbl, u64, i32],
        "marxb200_merge_events_end": [vp, vp],
        "marxb200_merge_download": [vp, vp, u64, vp],
        "marxb200_probe_d2h": [vp, u64, i32, i32, C.POINTER(dbl)],
        "marxb200_device_warmup": [i32],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    _LIB = lib
    return lib


class _PackedLayout(C.Structure):
    _fields_ = [("num_cols", C.c_uint32), ("n_rows", C.c_uint64), ("mask", C.c_uint64 * 32), ("file", (C.c_char * 16) * 32),
                ("type", C.c_char * 32), ("elem_size", C.c_uint32 * 32), ("offset", C.c_uint64 * 32)]


class _MergedLayout(C.Structure):
    _fields_ = [("num_cols", C.c_uint32), ("world", C.c_uint32), ("dst_rank", C.c_uint32), ("transport", C.c_uint32),
                ("n_rows", C.c_uint64), ("rows_of_rank", C.c_uint64 * 64), ("device_base", C.c_void_p),
                ("device_offset", C.c_uint64 * 32), ("mask", C.c_uint64 * 32), ("file", (C.c_char * 16) * 32),
                ("type", C.c_char * 32), ("elem_size", C.c_uint32 * 32), ("transfer_ms", C.c_double), ("copy_ms", C.c_double),
                ("nvlink_bytes", C.c_uint64)]


COMM_ID_BYTES = 128


def comm_unique_id():
    """rank 0: a fresh NCCL id (128 bytes) to hand to the other ranks (marxb200_comm_get_unique_id)"""
    lib = load_library()
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    if lib.marxb200_comm_get_unique_id(buf) != 0:
        raise MarxB200Error(lib.marxb200_last_error().decode(errors="replace"))
    return bytes(buf)


def shard_of(first_ray, n_total, rank, world):
    """(first ray, count) of rank's block in a sharded trace of rays [first_ray, first_ray + n_total)"""
    lib = load_library()
    f, n = C.c_uint64(), C.c_uint64()
    if lib.marxb200_shard_of(int(first_ray), int(n_total), int(rank), int(world), C.byref(f), C.byref(n)) != 0:
        raise MarxB200Error(lib.marxb200_last_error().decode(errors="replace"))
    return f.value, n.value


class _TallyAxis(C.Structure):
    _fields_ = [("column", C.c_int32), ("nbins", C.c_uint32), ("lo", C.c_double), ("hi", C.c_double)]


class Tally:
    """a device-resident histogram of the live event list; counts accumulate over batches until reset()"""

    def __init__(self, m, tid, shape):
        self._m, self.id, self.shape = m, tid, shape

    def accumulate(self):
        self._m._check(self._m._lib.marxb200_tally_accumulate(self._m._ctx, self.id))

    def reset(self):
        self._m._check(self._m._lib.marxb200_tally_reset(self._m._ctx, self.id))

    def read(self):
        out = np.zeros(self.shape, dtype=np.uint64)
        self._m._check(self._m._lib.marxb200_tally_read(self._m._ctx, self.id, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def allreduce(self):
        """sum the counters over all ranks of the context's communicator, in place (marxb200_tally_allreduce)"""
        self._m._check(self._m._lib.marxb200_tally_allreduce(self._m._ctx, self.id))

    def device_tensor(self):
        """the counters as a torch int64 tensor ALIASING the device buffer (for an in-place NCCL all-reduce)"""
        import torch
        ptr, n = C.c_void_p(), C.c_uint64()
        self._m._check(self._m._lib.marxb200_tally_device_ptr(self._m._ctx, self.id, C.byref(ptr), C.byref(n)))

        class _Alias:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<i8", "data": (int(ptr.value), False), "version": 2}
        return torch.as_tensor(_Alias(), device="cuda:%d" % self._m.device).view(self.shape)


class _Columns(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ["energy", "time", "xpos", "ypos", "zpos", "xcos", "ycos", "zcos", "chipx", "chipy", "pi",
                 "pha", "ccd", "order", "shell", "ray"]]


_COLUMN_DTYPES = {"energy": "<f8", "time": "<f8", "xpos": "<f8", "ypos": "<f8", "zpos": "<f8", "xcos": "<f8",
                  "ycos": "<f8", "zcos": "<f8", "chipx": "<f4", "chipy": "<f4", "pi": "<f4", "pha": "<i2",
                  "ccd": "i1", "order": "i1", "shell": "i1", "ray": "<u8"}


class MarxB200:
    """One context = one GPU = one photon list (the reference's single Marx_Photon_Type, marx.c:444)."""

    def __init__(self, calpack, device=0, seed=1, max_photons=1 << 20, stream=None):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        self.device = int(device)
        self._check(self._lib.marxb200_create(C.byref(self._ctx), int(device), int(seed)))
        try:
            if stream is not None:
                self._check(self._lib.marxb200_set_stream(self._ctx, C.c_void_p(int(stream))))
            path = calpack if os.path.exists(calpack) else caldata_path(calpack)
            self._check(self._lib.marxb200_load_calpack(self._ctx, path.encode()))
            self._check(self._lib.marxb200_alloc_photons(self._ctx, int(max_photons)))
        except Exception:
            self.close()
            raise
        self.max_photons = int(max_photons)

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise MarxB200Error(self._lib.marxb200_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.marxb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the module API ------------------------------------------------------------------------
    def set_compaction(self, on):
        self._check(self._lib.marxb200_set_compaction(self._ctx, 1 if on else 0))

    def create_photons(self, first_ray, n, time_base=-1.0):
        """marx_create_photons (source.c:268-384) for global ray indices [first_ray, first_ray+n)."""
        self._check(self._lib.marxb200_create_photons(self._ctx, int(first_ray), int(n), float(time_base)))

    def truncate_exposure(self, exposure_left):
        """ExposureTime handling of marx_create_photons (source.c:323-334); returns the number of rays kept."""
        n = C.c_uint64()
        self._check(self._lib.marxb200_truncate_exposure(self._ctx, float(exposure_left), C.byref(n)))
        return n.value

    def time_sums(self, first_ray, n):
        cap = n // 65536 + 2
        buf = (C.c_double * cap)()
        ns = C.c_uint64()
        self._check(self._lib.marxb200_time_sums(self._ctx, int(first_ray), int(n), buf, cap, C.byref(ns)))
        return np.frombuffer(buf, dtype=np.float64, count=ns.value).copy()

    def mirror_reflect(self):
        self._check(self._lib.marxb200_mirror_reflect(self._ctx))

    def grating_diffract(self):
        self._check(self._lib.marxb200_grating_diffract(self._ctx))

    def detect(self):
        self._check(self._lib.marxb200_detect(self._ctx))

    def restore_order(self):
        """put the live list back into arrival order (implicit in trace() and download*())"""
        self._check(self._lib.marxb200_restore_order(self._ctx))

    # ---- event tallies: device-resident histograms of the live list (include/marxb200.h) ----
    def tally_create(self, *axes):
        """axes: 1 or 2 tuples (column name, nbins, lo, hi); returns a Tally handle"""
        arr = (_TallyAxis * len(axes))()
        for k, (col, nbins, lo, hi) in enumerate(axes):
            arr[k].column, arr[k].nbins, arr[k].lo, arr[k].hi = TALLY_COLUMNS[col], int(nbins), float(lo), float(hi)
        tid = self._lib.marxb200_tally_create(self._ctx, C.cast(arr, C.c_void_p), len(axes))
        if tid < 0:
            raise MarxB200Error(self._lib.marxb200_last_error().decode())
        return Tally(self, tid, tuple(int(a[1]) for a in axes))

    def trace(self, first_ray, n, time_base=-1.0):
        """create -> mirror -> grating -> detect for one batch, device resident (marx.c:569, :240-273)."""
        self._check(self._lib.marxb200_trace_from(self._ctx, int(first_ray), int(n), float(time_base)))

    # -- multi-GPU: NCCL inside the library (include/marxb200.h "Multi-GPU") ----------------------------------------
    def comm_init(self, unique_id, rank, world):
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.marxb200_comm_init(self._ctx, buf, int(rank), int(world)))

    def comm_init_file(self, path, rank, world, timeout_s=120.0):
        self._check(self._lib.marxb200_comm_init_file(self._ctx, os.fsencode(path), int(rank), int(world), float(timeout_s)))

    def comm_info(self):
        r, w, v, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.marxb200_comm_info(self._ctx, C.byref(r), C.byref(w), C.byref(v), C.byref(t)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value,
                "merge_transport": "peer writes (CUDA IPC, copy engines)" if t.value == 1 else "ncclSend/ncclRecv"}

    def comm_destroy(self):
        self._check(self._lib.marxb200_comm_destroy(self._ctx))

    def trace_sharded(self, first_ray, n_total, time_base=-1.0):
        """collective: this rank traces its block of rays [first_ray, first_ray + n_total); -> (first ray, count) of the block"""
        f, n = C.c_uint64(), C.c_uint64()
        self._check(self._lib.marxb200_trace_sharded(self._ctx, int(first_ray), int(n_total), float(time_base), C.byref(f), C.byref(n)))
        return f.value, n.value

    def merge_events_begin(self, write_mask, total_time, max_rows_per_rank, dst_rank=0):
        self._check(self._lib.marxb200_merge_events_begin(self._ctx, int(write_mask), float(total_time), int(max_rows_per_rank), int(dst_rank)))

    def merge_events_end(self):
        lay = _MergedLayout()
        self._check(self._lib.marxb200_merge_events_end(self._ctx, C.byref(lay)))
        return {"n_rows": int(lay.n_rows), "rows_of_rank": [int(lay.rows_of_rank[r]) for r in range(lay.world)],
                "transport": int(lay.transport), "transfer_ms": float(lay.transfer_ms), "copy_ms": float(lay.copy_ms),
                "nvlink_bytes": int(lay.nvlink_bytes),
                "device_base": lay.device_base, "columns": [lay.file[j].value.decode() for j in range(lay.num_cols)]}

    def merge_download(self, host):
        """destination rank: the merged columns -> {file name: big-endian numpy view into `host`}"""
        lay = _PackedLayout()
        self._check(self._lib.marxb200_merge_download(self._ctx, host.ctypes.data_as(C.c_void_p), host.nbytes, C.byref(lay)))
        dts = {b"E": ">f4", b"I": ">i2", b"J": ">i4", b"A": "i1"}
        out = {}
        for j in range(lay.num_cols):
            dt = np.dtype(dts[lay.type[j:j + 1]])
            out[lay.file[j].value.decode()] = host[lay.offset[j]:lay.offset[j] + lay.n_rows * dt.itemsize].view(dt)
        return out

    def probe_d2h(self, nbytes, reps=8, write_combined=False, together=False):
        """device-to-host copy rate into pinned memory in GB/s (marxb200_probe_d2h); together: collective, all ranks start at once"""
        g = C.c_double()
        self._check(self._lib.marxb200_probe_d2h(self._ctx, int(nbytes), int(reps), (1 if write_combined else 0) | (2 if together else 0), C.byref(g)))
        return g.value

    KERNEL_CLASSES = ("k0_time_sums", "k0_time_scan", "k0_source", "k01_source_hrma", "k1_hrma<0>", "k1_hrma<1>",
                      "k1_hrma<2>", "k2_grating", "k3_detect", "order_restore", "level1",
                      "k1_hrma<B1>", "k1_hrma<B2C1>", "k1_hrma<C2>")

    def set_profiling(self, on):
        self._check(self._lib.marxb200_set_profiling(self._ctx, 1 if on else 0))

    def kernel_ms(self):
        """accumulated device milliseconds and launch counts per kernel class since the last call"""
        ms = (C.c_double * len(self.KERNEL_CLASSES))()
        nl = (C.c_uint64 * len(self.KERNEL_CLASSES))()
        self._check(self._lib.marxb200_get_kernel_ms(self._ctx, ms, nl))
        return {k: (float(ms[i]), int(nl[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    # -- results -------------------------------------------------------------------------------
    def counts(self):
        g, l, t = C.c_uint64(), C.c_uint64(), C.c_double()
        self._check(self._lib.marxb200_get_counts(self._ctx, C.byref(g), C.byref(l), C.byref(t)))
        return g.value, l.value, t.value

    def stage_counts(self):
        a = (C.c_uint64 * 4)()
        self._check(self._lib.marxb200_get_stage_counts(self._ctx, a))
        return [int(v) for v in a]

    def internal_counts(self):
        """[generated, after mirror, after grating, detected, after HRMA phase A, after HRMA phase B (B1 when the mirror stage
        runs as A | B1 | B2+C1 | C2), between the ACIS kernels, after HRMA B2+C1] of the last batch"""
        a = (C.c_uint64 * 8)()
        self._check(self._lib.marxb200_get_internal_counts(self._ctx, a))
        return [int(v) for v in a]

    def measure_fp64_peak(self):
        """measured FP64 peak of this GPU in TFLOP/s (DFMA chains; the FP64 roofline denominator)"""
        t = C.c_double()
        self._check(self._lib.marxb200_measure_fp64_peak(self._ctx, C.byref(t)))
        return t.value

    def launch_count(self):
        n = C.c_uint64()
        self._check(self._lib.marxb200_get_launch_count(self._ctx, C.byref(n)))
        return n.value

    def download(self, all_slots=False, out=None):
        """Live photons in arrival order as Marx_Photon_Attr_Type records (all slots if compaction is off)."""
        _, live, _ = self.counts()
        n = max(int(live), 1)
        if out is None:
            out = np.zeros(n, dtype=PHOTON_DTYPE)
        got = C.c_uint64()
        fn = self._lib.marxb200_download_all if all_slots else self._lib.marxb200_download
        self._check(fn(self._ctx, out.ctypes.data_as(C.c_void_p), len(out), C.byref(got)))
        return out[:got.value]

    def upload(self, photons, ray_ids=None, start_time=0.0):
        """Inject photons at a stage boundary (the reference's RAYFILE channel, s-rayfile.c:188-221); start_time is
        added to their arrival times (0: they are absolute)."""
        photons = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
        ids = None
        if ray_ids is not None:
            ids = np.ascontiguousarray(ray_ids, dtype=np.uint64)
        self._check(self._lib.marxb200_upload_from(self._ctx, photons.ctypes.data_as(C.c_void_p), len(photons),
                                                   ids.ctypes.data_as(C.c_void_p) if ids is not None else None, float(start_time)))

    def egress_begin(self, max_out):
        """snapshot the ordered live list on the device and return at once (see include/marxb200.h)"""
        self._check(self._lib.marxb200_egress_begin(self._ctx, int(max_out)))

    def egress_end(self, out):
        """out: dict name -> (pinned) numpy array; blocks until the snapshot has been copied; returns views"""
        cols = _Columns()
        for name, arr in out.items():
            setattr(cols, name, arr.ctypes.data)
        got = C.c_uint64()
        self._check(self._lib.marxb200_egress_end(self._ctx, C.byref(cols), C.byref(got)))
        return {k: v[:got.value] for k, v in out.items()}

    def egress_begin_packed(self, write_mask, total_time, max_out):
        """pack the selected columns into their MARX file images on the device and return at once (include/marxb200.h)"""
        self._check(self._lib.marxb200_egress_begin_packed(self._ctx, int(write_mask), float(total_time), int(max_out)))

    def egress_end_packed(self, host):
        """host: (pinned) uint8 numpy array; blocks until the packed columns have landed.  Returns {file name: big-endian
        numpy view into `host`}."""
        lay = _PackedLayout()
        self._check(self._lib.marxb200_egress_end_packed(self._ctx, host.ctypes.data_as(C.c_void_p), host.nbytes, C.byref(lay)))
        dts = {b"E": ">f4", b"I": ">i2", b"J": ">i4", b"A": "i1"}
        out = {}
        for j in range(lay.num_cols):
            dt = np.dtype(dts[lay.type[j:j + 1]])
            out[lay.file[j].value.decode()] = host[lay.offset[j]:lay.offset[j] + lay.n_rows * dt.itemsize].view(dt)
        return out

    def write_photons(self, directory, write_mask, open_mode, total_time):
        """marx_write_photons (marxio.c:403-476): create/append the column files of an output directory from the
        device-resident live list; write_mask uses the MARX_*_OK bits (HISTORY below)."""
        self._check(self._lib.marxb200_write_photons(self._ctx, os.fsencode(directory), int(write_mask),
                                                     1 if open_mode else 0, float(total_time)))

    def set_async_writer(self, n_threads):
        """background column-file writer for write_photons (0: back to synchronous writes)"""
        self._check(self._lib.marxb200_set_async_writer(self._ctx, int(n_threads)))

    def write_flush(self):
        self._check(self._lib.marxb200_write_flush(self._ctx))

    # -- Level-1 event transforms (marx2fits.c:3584-3943 on the device-resident list) -------------------------------
    def set_level1(self, desc):
        """desc: marx_b200.level1.Level1Desc (the values the stock marx2fits initialisation derives)"""
        self._level1_desc = desc              # keeps the arrays the descriptor points into alive during the call
        self._check(self._lib.marxb200_set_level1(self._ctx, desc.byref()))

    def level1_reset(self):
        self._check(self._lib.marxb200_level1_reset(self._ctx))

    def level1_transform(self, total_time=0.0):
        self._check(self._lib.marxb200_level1_transform(self._ctx, float(total_time)))

    def level1_download(self, names=None):
        from .level1 import alloc_columns
        _, live, _ = self.counts()
        cols, arrays = alloc_columns(int(live), names)
        got = C.c_uint64()
        self._check(self._lib.marxb200_level1_download(self._ctx, C.byref(cols), max(int(live), 1), C.byref(got)))
        return {k: v[:got.value] for k, v in arrays.items()}

    ASPSOL_COLUMNS = ("time", "ra", "dec", "roll", "q0", "q1", "q2", "q3")

    def aspsol_rows(self, desc, first_row, n, fits_rows=False):
        """Rows of the aspect-solution table (marxasp's row loop, marxasp.c:996-1027).  desc: the 21 doubles of
        marxb200_aspsol_desc.  -> (dict of the 8 f64 columns, the n*76-byte FITS row image or None, kernel milliseconds)"""
        d = np.ascontiguousarray(desc, dtype=np.float64)
        assert d.shape == (21,)
        cols = np.zeros((8, int(n)), dtype=np.float64)
        img = np.zeros(int(n) * 76, dtype=np.uint8) if fits_rows else None
        ms = C.c_double()
        self._check(self._lib.marxb200_aspsol_rows(self._ctx, d.ctypes.data_as(C.c_void_p), int(first_row), int(n),
                                                   cols.ctypes.data_as(C.c_void_p),
                                                   img.ctypes.data_as(C.c_void_p) if img is not None else None, C.byref(ms)))
        return {k: cols[j] for j, k in enumerate(self.ASPSOL_COLUMNS)}, img, ms.value

    PILEUP_DITHER = ("sky_ra", "sky_dec", "sky_roll", "det_dy", "det_dz", "det_theta")

    def pileup(self, cols, alpha, frame_time, seed, max_out=None, out=None):
        """ACIS pile-up on the event columns of a simulation (marxpileup's frame loop, marxpileup.c:1121-1213).  cols: dict with
        ccd (i8), x, y, t, benergy (f32) and optionally the six dither columns, in file order; frame_time = FrameTime +
        FrameTransferTime.  -> (dict of the output columns write_event :622-666 writes, milliseconds of the device kernels)"""
        n = len(cols["t"])
        keep = {"ccd": np.ascontiguousarray(cols["ccd"], dtype=np.int8)}     # no copy when the caller's (pinned) array already fits
        for k in ("x", "y", "t", "benergy") + tuple(d for d in self.PILEUP_DITHER if d in cols):
            keep[k] = np.ascontiguousarray(cols[k], dtype=np.float32)
            assert len(keep[k]) == n
        ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
        pin = (C.c_void_p * 11)(*[ptr(keep.get(k)) for k in ("ccd", "x", "y", "t", "benergy") + self.PILEUP_DITHER])
        cap = n if max_out is None else int(max_out)
        if out is None:
            out = {"ccd": np.zeros(cap, np.int8), "x": np.zeros(cap, np.float32), "y": np.zeros(cap, np.float32), "t": np.zeros(cap, np.float32),
                   "benergy": np.zeros(cap, np.float32), "frame": np.zeros(cap, np.int32), "nphotons": np.zeros(cap, np.int16),
                   "pha": np.zeros(cap, np.int16)}
            for d in self.PILEUP_DITHER:
                if d in keep:
                    out[d] = np.zeros(cap, np.float32)
        pout = (C.c_void_p * 14)(*[ptr(out.get(k)) for k in ("ccd", "x", "y", "t", "benergy", "frame", "nphotons", "pha") + self.PILEUP_DITHER])
        got, ms = C.c_uint64(), C.c_double()
        self._check(self._lib.marxb200_pileup_run(self._ctx, n, pin, float(alpha), float(frame_time), int(seed), cap, pout,
                                                  C.byref(got), C.byref(ms)))
        return {k: v[:got.value] for k, v in out.items()}, ms.value

    def pileup_events(self, total_time, alpha, frame_time, seed, max_out=None, out=None):
        """marxb200_pileup_events: the pile-up model on the event list the detector stage left on the device (no column files in
        between).  -> (dict of output columns incl. the six dither columns, milliseconds of the device kernels)"""
        _, live, _ = self.counts()
        cap = int(live) if max_out is None else int(max_out)
        cap = max(cap, 1)
        if out is None:
            out = {"ccd": np.zeros(cap, np.int8), "x": np.zeros(cap, np.float32), "y": np.zeros(cap, np.float32), "t": np.zeros(cap, np.float32),
                   "benergy": np.zeros(cap, np.float32), "frame": np.zeros(cap, np.int32), "nphotons": np.zeros(cap, np.int16),
                   "pha": np.zeros(cap, np.int16)}
            for d in self.PILEUP_DITHER:
                out[d] = np.zeros(cap, np.float32)
        ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
        pout = (C.c_void_p * 14)(*[ptr(out.get(k)) for k in ("ccd", "x", "y", "t", "benergy", "frame", "nphotons", "pha") + self.PILEUP_DITHER])
        got, ms = C.c_uint64(), C.c_double()
        self._check(self._lib.marxb200_pileup_events(self._ctx, float(total_time), float(alpha), float(frame_time), int(seed), cap, pout,
                                                     C.byref(got), C.byref(ms)))
        return {k: v[:got.value] for k, v in out.items()}, ms.value

    def download_columns(self, names=("energy", "time", "chipx", "chipy", "pha", "ccd", "order", "ray"), out=None):
        _, live, _ = self.counts()
        n = max(int(live), 1)
        cols = _Columns()
        arrays = {}
        for name in names:
            if out is not None and name in out:
                arr = out[name]
            else:
                arr = np.empty(n, dtype=_COLUMN_DTYPES[name])
            arrays[name] = arr
            setattr(cols, name, arr.ctypes.data)
        got = C.c_uint64()
        cap = min(len(a) for a in arrays.values()) if arrays else 0
        self._check(self._lib.marxb200_download_columns(self._ctx, C.byref(cols), cap, C.byref(got)))
        return {k: v[:got.value] for k, v in arrays.items()}
