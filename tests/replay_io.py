"""Reader for the replay dumps written by oracle/_ref/marx_replay (oracle/ref/replay_harness.c)."""
import numpy as np

from marx_b200.api import PHOTON_DTYPE

REPLAY_DTYPE = np.dtype([("st", PHOTON_DTYPE, 4), ("draws", "<u4", 4), ("start", "<f8")])


def read_replay(path):
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[:8].tobytes() == b"MRXRPLY1", "not a replay dump"
    nrays, seed, first = np.frombuffer(raw, "<u8", 3, 8)
    nstages, recsize = np.frombuffer(raw, "<u4", 2, 32)
    modules = np.frombuffer(raw, "<i4", 4, 40)
    assert nstages == 4 and recsize == PHOTON_DTYPE.itemsize
    recs = raw[56:].view(REPLAY_DTYPE)
    assert len(recs) == nrays
    return dict(nrays=int(nrays), seed=int(seed), first_ray=int(first), modules=modules.copy()), recs
