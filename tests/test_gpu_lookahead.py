"""The arrival-time pre-pass of the NEXT contiguous batch runs behind the batch being traced (marxb200_trace, DESIGN.md section 4
"Work behind the batch being traced").  Whatever the caller asks for next -- the predicted batch, other rays, another batch size, the same
rays after marxb200_set_source -- the events must be those of a context that never looks ahead (MARXB200_LOOKAHEAD=0), bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
COLS = ("energy", "time", "xpos", "ypos", "zpos", "xcos", "ycos", "zcos", "chipx", "chipy", "pi", "pha", "ccd", "order", "shell", "ray")
N = 1 << 18


def _context(calpack, lookahead):
    import marx_b200
    old = os.environ.get("MARXB200_LOOKAHEAD")
    os.environ["MARXB200_LOOKAHEAD"] = "1" if lookahead else "0"          # read by marxb200_create
    try:
        return marx_b200.MarxB200(calpack, seed=4242, max_photons=N)
    finally:
        if old is None:
            del os.environ["MARXB200_LOOKAHEAD"]
        else:
            os.environ["MARXB200_LOOKAHEAD"] = old


def _reload(m, calpack):
    from marx_b200.api import caldata_path
    m._check(m._lib.marxb200_load_calpack(m._ctx, caldata_path(calpack).encode()))


def _run(lookahead, script):
    """script: list of ("trace", first, n, time_base) / ("load", calpack); -> the event lists after every trace"""
    out = []
    with _context("c2_hetg_acis_s", lookahead) as m:
        for step in script:
            if step[0] == "load":
                _reload(m, step[1])
            else:
                m.trace(step[1], step[2], step[3])
                out.append((m.download_columns(COLS), m.counts()))
    return out


def _same(a, b, skip=()):
    assert len(a) == len(b)
    for k, ((ea, ca), (eb, cb)) in enumerate(zip(a, b)):
        assert ca == cb, (k, ca, cb)
        for c in COLS:
            if c not in skip:
                assert ea[c].tobytes() == eb[c].tobytes(), (k, c)


@pytest.mark.parametrize("script", [
    # the predicted batch (hit), twice, continuing the running time
    [("trace", 0, N, 0.0), ("trace", N, N, -1.0), ("trace", 2 * N, N, -1.0)],
    # other rays than predicted, then a hit again, then another batch size and an explicit time base
    [("trace", 0, N, 0.0), ("trace", 5 * N, N, -1.0), ("trace", 6 * N, N, -1.0), ("trace", 7 * N, N // 2 + 77, 1.0e6), ("trace", 0, N, 3.5)],
    # the predicted rays, but of ANOTHER source (BETA instead of POINT: another draw of the time increment): the sums made ahead are stale
    [("trace", 0, N, 0.0), ("load", "c4_beta_acis_i"), ("trace", N, N, -1.0), ("trace", 2 * N, N, -1.0)],
], ids=["hits", "misses", "source_changed"])
def test_events_do_not_depend_on_the_look_ahead(script):
    # GratingType=NONE (c4) never writes the diffraction order: after a HETG batch in the same context the column holds what that
    # batch left in whatever rows the compaction happened to use (marx_write_photons does not emit it either, marxio.c:403-476)
    skip = ("order",) if any(step[0] == "load" for step in script) else ()
    _same(_run(True, script), _run(False, script), skip)


def _egress_run(prepack, total_times):
    """three batches, each followed by the pipelined packed egress with the same columns; -> the packed column images per batch"""
    import marx_b200
    from marx_b200 import HISTORY
    mask = sum(HISTORY[k] for k in ("ENERGY", "TIME", "X_VECTOR", "P_VECTOR", "DET_NUM", "DET_PIXEL", "MIRROR_SHELL", "PULSEHEIGHT",
                                    "ORDER", "PI", "SKY_DITHER", "TAG"))
    old = os.environ.get("MARXB200_PREPACK")
    os.environ["MARXB200_PREPACK"] = "1" if prepack else "0"              # read when a begin call arms the next batch
    out = []
    try:
        with marx_b200.MarxB200("c2_hetg_acis_s", seed=99, max_photons=N) as m:
            host = np.zeros(N // 4 * 96 + 4096, dtype=np.uint8)
            launches = []
            for k, tt in enumerate(total_times):
                m.trace(k * N, N)
                before = m.launch_count()
                m.egress_begin_packed(mask, tt, N // 4)
                launches.append(m.launch_count() - before)
                cols = m.egress_end_packed(host)
                out.append({name: v.copy() for name, v in cols.items()})
    finally:
        if old is None:
            del os.environ["MARXB200_PREPACK"]
        else:
            os.environ["MARXB200_PREPACK"] = old
    return out, launches


def test_file_images_written_by_the_order_restoration_equal_the_conversion_kernel():
    """pre-pack: from the second batch on the packed images come from order_gather<true>; a changed time offset
    falls back to the conversion kernel for that batch"""
    times = [0.0, 0.0, 0.0, 12.5, 12.5]
    a, la = _egress_run(True, times)
    b, lb = _egress_run(False, times)
    assert la == [1, 0, 0, 1, 0] and lb == [1, 1, 1, 1, 1], (la, lb)      # conversion kernels launched by the begin calls
    for k, (x, y) in enumerate(zip(a, b)):
        assert set(x) == set(y) and len(x) >= 12
        for name in x:
            assert x[name].tobytes() == y[name].tobytes(), (k, name)
