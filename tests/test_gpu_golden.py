"""GPU parity against the committed golden fixtures (reference run through oracle/_ref/marx_replay)."""
import os

import numpy as np
import pytest

from tests.parity import assert_stage_ok, compare_stage, rel

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + "_replay.npz"))
    return int(z["seed"]), int(z["first_ray"]), z["stages"], z["start_time"]


@pytest.mark.parametrize("config", ["c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s", "c4_beta_acis_i", "c4_image_acis_i", "c1_line_acis_s", "c3_hrc_i"])
def test_replay_parity_in_place(config):
    """compaction off: every ray keeps its slot, so each stage is compared ray by ray, dead rays included."""
    import marx_b200
    seed, first, stages, start = _load(config)
    n = len(stages)
    with marx_b200.MarxB200(config, seed=seed, max_photons=n) as m:
        m.set_compaction(False)
        m.create_photons(first, n, time_base=0.0)
        got = m.download(all_slots=True)
        assert len(got) == n
        f = compare_stage(got, stages[:, 0], 0)
        assert_stage_ok(f, 0)
        # absolute arrival times: reference start_time + arrival_time (batch size 1) vs the device scan
        t_ref = start + stages[:, 0]["arrival_time"]
        assert rel(got["arrival_time"], t_ref).max() <= 1e-12
        assert (got["tag"] == (first + np.arange(n)).astype(np.uint32)).all()
        for stage, call in ((1, m.mirror_reflect), (2, m.grating_diffract), (3, m.detect)):
            call()
            got = m.download(all_slots=True)
            f = compare_stage(got, stages[:, stage], stage)
            print(config, "stage", stage, f)
            assert_stage_ok(f, stage)


@pytest.mark.parametrize("config", ["c1_acis_s", "c2_hetg_acis_s", "c3_letg_hrc_s", "c4_beta_acis_i", "c4_image_acis_i", "c1_line_acis_s", "c3_hrc_i"])
def test_replay_parity_compacted(config):
    """the product path: fused compaction; survivors must be the reference's survivors in arrival order."""
    import marx_b200
    seed, first, stages, start = _load(config)
    n = len(stages)
    with marx_b200.MarxB200(config, seed=seed, max_photons=n) as m:
        m.create_photons(first, n, time_base=0.0)
        for stage, call in ((1, m.mirror_reflect), (2, m.grating_diffract), (3, m.detect)):
            call()
            got = m.download()
            ref = stages[:, stage]
            ref = ref[(ref["flags"] & 0xFF) == 0]
            assert len(got) == len(ref), (stage, len(got), len(ref))
            assert (got["tag"] == ref["tag"]).all()
            f = compare_stage(got, ref, stage)
            assert_stage_ok(f, stage)
        counts = m.stage_counts()
        assert counts[0] == n and counts[3] == len(got)


def test_trace_is_batch_invariant():
    """counter-based draws: tracing [0,N) in one batch or in two gives the same events (times continue)."""
    import marx_b200
    n = 8192
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=3, max_photons=n) as m:
        m.create_photons(0, n, time_base=0.0)
        m.mirror_reflect(); m.grating_diffract(); m.detect()
        one = m.download().copy()
        t_end = m.counts()[2]
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=3, max_photons=n) as m:
        parts = []
        m.create_photons(0, 4096, time_base=0.0)
        t0 = 0.0
        m.mirror_reflect(); m.grating_diffract(); m.detect()
        a = m.download().copy(); t_mid = m.counts()[2]
        m.trace(4096, 4096)
        b = m.download().copy()
        assert abs(m.counts()[2] - t_end) <= 1e-9 * t_end
    two_tags = np.concatenate([a["tag"], b["tag"]])
    assert (two_tags == one["tag"]).all()
    for k in ("energy", "x", "p", "pulse_height", "ccd_num", "order", "y_pixel", "z_pixel"):
        assert (np.concatenate([a[k], b[k]]) == one[k]).all(), k
    t_two = np.concatenate([a["arrival_time"], b["arrival_time"] + t_mid])
    assert rel(t_two, one["arrival_time"]).max() <= 1e-12
