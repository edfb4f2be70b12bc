"""Event tallies (marxb200_tally_*): device-resident histograms of the live list, checked bin by bin against the same
binning of the downloaded events (exact integer counts), accumulated over batches, on the compacted and the in-place
list, 1-D and 2-D, small (shared-memory) and large (global-atomics) shapes."""
import numpy as np
import pytest

import marx_b200

pytestmark = pytest.mark.gpu


def np_tally(cols, axes):
    """the documented binning: floor((v - lo) * nbins / (hi - lo)), dropped outside [0, nbins) on any axis"""
    ok = np.ones(len(next(iter(cols.values()))), dtype=bool)
    idx = []
    for col, nbins, lo, hi in axes:
        v = cols[col].astype(np.float64)
        f = np.floor((v - lo) * (nbins / (hi - lo)))
        ok &= (f >= 0) & (f < nbins)
        idx.append(np.where(ok, f, 0).astype(np.int64))
    out = np.zeros([a[1] for a in axes], dtype=np.uint64)
    np.add.at(out, tuple(i[ok] for i in idx), 1)
    return out


SPECS = [
    (("order", 23, -11, 12),),
    (("pha", 4096, 0, 4096),),
    (("energy", 770, 0.3, 8.0),),
    (("pi", 1000, 0.0, 12.0),),
    (("ccd", 10, 0, 10), ("order", 7, -3, 4)),
    (("shell", 8, 0, 8),),
    (("chipx", 256, 0, 1024), ("chipy", 256, 0, 1024)),          # 65536 bins: the global-atomics path
    (("time", 100, 0.0, 3.0e6),),
    (("ypos", 128, -40.0, 40.0), ("zpos", 64, -5.0, 5.0)),
]
NAMES = ("energy", "time", "pha", "pi", "order", "ccd", "shell", "chipx", "chipy", "ypos", "zpos")


@pytest.mark.parametrize("compact", [True, False])
def test_tallies_equal_binned_events_over_batches(compact):
    n, batches = 1 << 20, 3
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=21, max_photons=n) as m:
        m.set_compaction(compact)
        # the AoS records of the in-place download carry batch-relative times (marx.h arrival_time): the absolute TIME
        # axis is compared on the compacted list only, whose column download is absolute like the tally
        specs = [sp for sp in SPECS if compact or all(a[0] != "time" for a in sp)]
        tallies = [m.tally_create(*spec) for spec in specs]
        n_events = 0
        want = [np.zeros(t.shape, dtype=np.uint64) for t in tallies]
        for b in range(batches):
            m.create_photons(b * n, n, time_base=(0.0 if b == 0 else -1.0))
            m.mirror_reflect(); m.grating_diffract(); m.detect()
            for t in tallies:
                t.accumulate()
            if compact:
                cols = m.download_columns(NAMES)
            else:
                ph = m.download(all_slots=True)
                ph = ph[(ph["flags"] & 0xFF) == 0]
                cols = {"energy": ph["energy"], "time": ph["arrival_time"], "pha": ph["pulse_height"], "pi": ph["pi"],
                        "order": ph["order"], "ccd": ph["ccd_num"], "shell": ph["mirror_shell"], "chipx": ph["y_pixel"],
                        "chipy": ph["z_pixel"], "ypos": ph["x"][:, 1], "zpos": ph["x"][:, 2]}
            n_events += len(cols["pha"])
            for k, spec in enumerate(specs):
                want[k] += np_tally(cols, spec)
        for k, t in enumerate(tallies):
            got = t.read()
            assert got.sum() > 0 and (got == want[k]).all(), specs[k]
        # every detected event has an order in -11..11 and a PHA in 0..4095: those tallies count them all
        assert tallies[0].read().sum() == tallies[1].read().sum() == n_events
        tallies[0].reset()
        assert tallies[0].read().sum() == 0


def test_tally_after_an_earlier_stage_and_device_alias():
    """tallies bin whatever list is current (here: after the mirror), and the torch alias of the device buffer sees
    the same counters (this is the tensor NCCL all-reduces in place, marx_b200.dist.allreduce_tally)"""
    n = 1 << 19
    with marx_b200.MarxB200("c1_acis_s", seed=3, max_photons=n) as m:
        t = m.tally_create(("shell", 8, 0, 8))
        m.create_photons(0, n, time_base=0.0)
        m.mirror_reflect()
        t.accumulate()
        ph = m.download()
        want = np.bincount(ph["mirror_shell"], minlength=8).astype(np.uint64)
        assert (t.read() == want).all()
        assert set(np.nonzero(want)[0]) == {0, 1, 2, 3}           # mirror_shell is the index of the HRMA shell (1, 3, 4, 6)
        alias = t.device_tensor()
        assert (alias.cpu().numpy().astype(np.uint64) == want).all()
        alias += 1                                                 # in place on the library's buffer
        import torch
        torch.cuda.synchronize()
        assert (t.read() == want + 1).all()


def test_tally_argument_errors():
    with marx_b200.MarxB200("c1_acis_s", seed=1, max_photons=1024) as m:
        with pytest.raises(marx_b200.MarxB200Error):
            m.tally_create(("pha", 0, 0, 10))
        with pytest.raises(marx_b200.MarxB200Error):
            m.tally_create(("pha", 10, 5, 5))
        t = m.tally_create(("pha", 16, 0, 4096))
        with pytest.raises(marx_b200.MarxB200Error):
            t.accumulate()                                         # no photons yet
