"""marx_b200 -- B200-native implementation of MARX's per-photon ray-trace path.

The product is ``libmarxb200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/marxb200.h``).  This package is only the thin Python host mirror used by the tests and
``bench.py``: ctypes bindings plus numpy views of the reference's photon record.  There is no CPU
fallback: importing works without a GPU, every compute call raises if the library or a CUDA device is
missing.
"""
from .api import (MarxB200, MarxB200Error, PHOTON_DTYPE, caldata_path, lib_path, load_library,
                  STAGE_SOURCE, STAGE_MIRROR, STAGE_GRATING, STAGE_DETECTOR, EXPORTED_SYMBOLS, HISTORY, read_marx_column)

__all__ = ["MarxB200", "MarxB200Error", "PHOTON_DTYPE", "caldata_path", "lib_path", "load_library",
           "STAGE_SOURCE", "STAGE_MIRROR", "STAGE_GRATING", "STAGE_DETECTOR", "EXPORTED_SYMBOLS", "HISTORY",
           "read_marx_column"]
