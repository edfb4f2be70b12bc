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
