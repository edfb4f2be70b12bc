"""ctypes access to the test-only CPU restatement (oracle/liboracle.so).  Import from tests only."""
import ctypes as C
import os

import numpy as np

from marx_b200.api import PHOTON_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _LIB = C.CDLL(path)
        _LIB.oracle_open.restype = C.c_void_p
        _LIB.oracle_open.argtypes = [C.c_char_p, C.c_uint64]
        _LIB.oracle_close.argtypes = [C.c_void_p]
        _LIB.oracle_last_error.restype = C.c_char_p
        _LIB.oracle_trace.restype = C.c_long
        _LIB.oracle_trace.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)] + [C.c_void_p] * 4
        _LIB.oracle_last_generated.restype = C.c_uint64
        _LIB.oracle_last_generated.argtypes = [C.c_void_p]
    return _LIB


class Oracle:
    def __init__(self, calpack, seed):
        from marx_b200.api import caldata_path
        path = calpack if os.path.exists(calpack) else caldata_path(calpack)
        self._h = lib().oracle_open(path.encode(), int(seed))
        if not self._h:
            raise RuntimeError(lib().oracle_last_error().decode())

    def trace(self, first_ray, n, time_base=0.0):
        """returns (stages[4][n] records with ABSOLUTE arrival_time, end_time, n_detected)"""
        st = np.zeros((4, n), dtype=PHOTON_DTYPE)
        tb = C.c_double(time_base)
        nd = lib().oracle_trace(self._h, int(first_ray), int(n), C.byref(tb),
                                *[st[s].ctypes.data_as(C.c_void_p) for s in range(4)])
        if nd < 0:
            raise RuntimeError("oracle_trace failed")
        self.last_generated = int(lib().oracle_last_generated(self._h))   # < n: the ASPSOL file ended inside the batch
        return st, tb.value, int(nd)

    def close(self):
        if self._h:
            lib().oracle_close(self._h)
            self._h = None

    def __del__(self):
        self.close()
