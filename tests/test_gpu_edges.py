"""Edge cases of the C ABI on the GPU: empty and ragged batches, capacity and call-order errors, the ExposureTime cut,
empty event lists through the writers.  The reference's analogues: `marx_create_photons` with num = 0 / an exposure
that ends inside a batch (source.c:268-384), `marx_write_photons` on a batch without survivors (marxio.c:403-476)."""
import os

import numpy as np
import pytest

import marx_b200
from marx_b200 import HISTORY, MarxB200Error, read_marx_column

pytestmark = pytest.mark.gpu


def _trace(m, first, n):
    m.create_photons(first, n, time_base=0.0)
    m.mirror_reflect(); m.grating_diffract(); m.detect()
    return m.download().copy()


@pytest.mark.parametrize("n", [1, 31, 33, 255, 257, 65535, 65537])
def test_ragged_batches_are_prefixes_of_a_full_batch(n):
    """a batch of n rays (n not a multiple of the 32-ray warp tile, the 256-ray time tile or the 65536-ray super-tile)
    yields exactly the events of rays [0, n) of a larger batch, staged and fused"""
    cap = 1 << 17
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=3, max_photons=cap) as m:
        full = _trace(m, 0, cap)
        part = _trace(m, 0, n)
        m.trace(0, n, 0.0)
        fused = m.download().copy()
        assert m.stage_counts()[0] == n
    want = full[full["tag"] < n]
    for got in (part, fused):
        assert len(got) == len(want)
        for k in ("tag", "energy", "x", "p", "arrival_time", "pulse_height", "ccd_num", "order", "y_pixel", "z_pixel", "pi"):
            assert (got[k] == want[k]).all(), (n, k)


def test_empty_batch_and_empty_event_list(tmp_path):
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=3, max_photons=4096) as m:
        m.trace(0, 0)
        assert m.stage_counts() == [0, 0, 0, 0] and len(m.download()) == 0
        m.create_photons(0, 0)
        m.mirror_reflect(); m.grating_diffract(); m.detect()
        assert m.counts()[1] == 0
        # one ray that certainly dies somewhere: the writers must still produce well-formed (possibly empty) files
        os.makedirs(tmp_path / "o")
        mask = HISTORY["ENERGY"] | HISTORY["TIME"] | HISTORY["PULSEHEIGHT"] | HISTORY["DET_NUM"]
        m.trace(0, 0)
        m.write_photons(tmp_path / "o", mask, True, 0.0)
        for f in ("energy.dat", "time.dat", "pha.dat", "detector.dat"):
            name, data = read_marx_column(tmp_path / "o" / f)
            assert len(data) == 0 and os.path.getsize(tmp_path / "o" / f) == 32
        m.trace(0, 4096)
        m.write_photons(tmp_path / "o", mask, False, 0.0)
        n_live = m.counts()[1]
        assert len(read_marx_column(tmp_path / "o" / "energy.dat")[1]) == n_live > 0
        m.upload(np.zeros(0, dtype=marx_b200.PHOTON_DTYPE))
        assert len(m.download()) == 0


def test_errors_are_reported_not_guessed():
    with marx_b200.MarxB200("c1_acis_s", seed=1, max_photons=1024) as m:
        with pytest.raises(MarxB200Error, match="exceeds the allocated capacity"):
            m.create_photons(0, 1025)
        with pytest.raises(MarxB200Error, match="no photons"):
            m.mirror_reflect()
        with pytest.raises(MarxB200Error, match="directly after"):
            m.trace(0, 512)
            m.truncate_exposure(1.0)
    with pytest.raises(MarxB200Error, match="calpack"):
        marx_b200.MarxB200(os.devnull)


def test_exposure_cut_semantics():
    """source.c:323-334: keep rays up to and INCLUDING the first one whose arrival time reaches the exposure"""
    n = 1 << 16
    with marx_b200.MarxB200("c1_acis_s", seed=9, max_photons=n) as m:
        m.set_compaction(False)
        m.create_photons(0, n, time_base=0.0)
        t = m.download(all_slots=True)["arrival_time"].copy()
        assert (np.diff(t) >= 0).all()
        for expo in (t[0] * 0.5, t[100], 0.5 * (t[1000] + t[1001]), t[-1], t[-1] * 2.0):
            m.create_photons(0, n, time_base=0.0)
            kept = m.truncate_exposure(expo)
            idx = np.nonzero(t >= expo)[0]
            want = (idx[0] + 1) if len(idx) else n
            assert kept == want, (expo, kept, want)
            gen, live, t_end = m.counts()
            assert gen == kept == live and t_end == t[kept - 1]
            m.mirror_reflect(); m.detect()
            tags = m.download(all_slots=True)["tag"]
            assert len(tags) == kept


def test_sixty_four_bit_seed_and_ray_index():
    """RandomSeed is an unsigned long in the reference (marx.c:846-849); ray indices beyond 2^32 select new streams"""
    n = 1 << 16
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=(1 << 40) + 5, max_photons=n) as m:
        a = _trace(m, 0, n)
        b = _trace(m, 1 << 36, n)
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=5, max_photons=n) as m:
        c = _trace(m, 0, n)
    assert len(a) and len(b) and len(c)
    assert not (len(a) == len(c) and (a["pulse_height"] == c["pulse_height"]).all())      # high seed bits matter
    assert not (len(a) == len(b) and (a["pulse_height"] == b["pulse_height"]).all())      # high ray-index bits matter
    assert (b["tag"] == (b["tag"].astype(np.uint64) & 0xFFFFFFFF)).all()


def test_packed_pipelined_egress_equals_the_written_files(tmp_path):
    """marxb200_egress_begin_packed/_end_packed (copy overlapped with the next batch) lands exactly the bytes that
    marxb200_write_photons appends to the column files"""
    n = 1 << 19
    mask = sum(HISTORY[k] for k in ("ENERGY", "TIME", "X_VECTOR", "P_VECTOR", "PULSEHEIGHT", "PI", "DET_PIXEL", "DET_NUM",
                                    "MIRROR_SHELL", "ORDER", "SKY_DITHER", "DET_DITHER", "TAG"))
    os.makedirs(tmp_path / "o")
    host = [np.zeros(n // 4 * 120, dtype=np.uint8) for _ in range(2)]
    got = []
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=21, max_photons=n) as m:
        total = 0.0
        for b in range(3):
            m.trace(b * n, n)
            m.write_photons(tmp_path / "o", mask, b == 0, total)
            if b > 0:
                got.append({k: v.copy() for k, v in m.egress_end_packed(host[(b - 1) & 1]).items()})
            m.egress_begin_packed(mask, total, n // 4)
            total = m.counts()[2]
        got.append({k: v.copy() for k, v in m.egress_end_packed(host[0]).items()})
    assert len(got[0]) == 22
    for f in got[0]:
        name, data = read_marx_column(tmp_path / "o" / f)
        cat = np.concatenate([g[f] for g in got])
        assert len(cat) == len(data) > 0 and (cat.astype(data.dtype) == data).all(), f
