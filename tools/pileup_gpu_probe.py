"""First-contact probe for marxb200_pileup_run on a GPU box: the three committed fixtures against the stock output, no torch import
(ctypes + numpy only, a few seconds).  Writes gpurun_out/pileup_probe.txt."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import marx_b200  # noqa: E402
from tests import pileup_lib as P  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
lines = []
for name in P.CASES:
    cols, ref, seed = P.load_golden(name)
    alpha, ft = P.pileup_params(P.CASES[name][1])
    t0 = time.time()
    with marx_b200.MarxB200(P.CASES[name][2], seed=1, max_photons=1024) as m:
        got, ms = m.pileup(cols, alpha, ft, seed)
    bad = [k for k in ref if got[k].tobytes() != ref[k].tobytes()]
    lines.append("%s: %d events -> %d rows (stock %d), kernels %.3f ms, wall %.2f s, differing columns: %s"
                 % (name, len(cols["t"]), len(got["t"]), len(ref["t"]), ms, time.time() - t0, bad or "none"))
    print(lines[-1], flush=True)
    open(os.path.join(ROOT, "gpurun_out", "pileup_probe.txt"), "w").write("\n".join(lines) + "\n")

if "--full" in sys.argv:
    import numpy as np

    def note(s):
        lines.append(s)
        print(s, flush=True)
        open(os.path.join(ROOT, "gpurun_out", "pileup_probe.txt"), "w").write("\n".join(lines) + "\n")

    # the workload of bench.py's "pileup" leg
    n = 1 << 22
    r = np.random.default_rng(1)
    alpha, ft, rate = 0.5, 3.241, 10.0
    cols = {"ccd": np.full(n, 7, np.int8), "t": np.cumsum(r.exponential(1.0 / rate, n)).astype(np.float32),
            "x": (512.0 + r.normal(0.0, 3.0, n)).astype(np.float32), "y": (300.0 + r.normal(0.0, 3.0, n)).astype(np.float32),
            "benergy": r.uniform(0.4, 7.0, n).astype(np.float32)}
    for k in P.DITHER:
        cols[k] = r.normal(0.0, 1e-3, n).astype(np.float32)
    with marx_b200.MarxB200("c2_hetg_acis_s", seed=1, max_photons=1024) as m:
        m.pileup(cols, alpha, ft, 1)
        ms, wall = [], []
        for _ in range(5):
            t0 = time.time()
            got, k = m.pileup(cols, alpha, ft, 1)
            wall.append((time.time() - t0) * 1e3)
            ms.append(k)
        t0 = time.time()
        ref = P.oracle_pileup(cols, ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"], "c2_hetg_acis_s", 1)
        dt = time.time() - t0
        bad = [k for k in ref if got[k].tobytes() != ref[k].tobytes()]
        note("bench workload: %d events -> %d rows, kernels median %.3f ms (%s), call with host columns median %.1f ms, oracle one core %.2f s, "
             "differing columns: %s" % (n, len(got["t"]), float(np.median(ms)), " ".join("%.3f" % v for v in ms), float(np.median(wall)), dt, bad or "none"))
        from tests.test_gpu_zz_pileup import _synthetic
        for nn, rate, alpha, ft, spot, ccd in [(200000, 40.0, 0.5, 3.241, 6.0, (7,)), (300000, 2.0, 0.9, 3.2, 1.5, (7,)),
                                               (100000, 500.0, 0.2, 0.4, 30.0, (5, 6, 7, 8)), (50000, 3000.0, 1.0, 3.2, 40.0, (7,))]:
            cols = _synthetic(nn, rate, 1234 + nn, spot=spot, ccd=ccd)
            ref = P.oracle_pileup(cols, ["Alpha=%r" % alpha, "FrameTime=%r" % ft, "FrameTransferTime=0.0"], "c1_acis_s", 77)
            got, k = m.pileup(cols, alpha, ft, 77)
            bad = [c for c in ref if got[c].tobytes() != ref[c].tobytes()]
            note("synthetic %d events at %.0f/s: %d rows (oracle %d), kernels %.3f ms, differing columns: %s"
                 % (nn, rate, len(got["t"]), len(ref["t"]), k, bad or "none"))

if "--edges" in sys.argv:
    import numpy as np
    cols, ref, seed = P.load_golden("pileup_acis_s_moderate")
    alpha, ft = P.pileup_params(P.CASES["pileup_acis_s_moderate"][1])
    out = []

    def refusal(m, c, a=alpha, **kw):
        try:
            m.pileup(c, a, ft, seed, **kw)
            return "NO ERROR"
        except marx_b200.MarxB200Error as e:
            return "refused: " + str(e)

    with marx_b200.MarxB200("c1_acis_s", seed=1, max_photons=1024) as m:
        out.append("empty: %d rows" % len(m.pileup({k: v[:0] for k, v in cols.items()}, alpha, ft, seed)[0]["t"]))
        one = m.pileup({k: v[:1] for k, v in cols.items()}, alpha, ft, seed)[0]
        out.append("one event: %d rows, nphotons %s, x same %s" % (len(one["t"]), one["nphotons"], one["x"][0] == cols["x"][0]))
        piled = m.pileup({k: np.repeat(v[:1], 5) for k, v in cols.items()}, 1.0, ft, seed)[0]
        out.append("5 identical: %d rows, nphotons %s, benergy %s vs %s" % (len(piled["t"]), piled["nphotons"], piled["benergy"], 5 * cols["benergy"][0]))
        out.append("max_out - 1: " + refusal(m, cols, max_out=len(ref["t"]) - 1))
        exact = m.pileup(cols, alpha, ft, seed, max_out=len(ref["t"]))[0]
        out.append("exact capacity identical: %s" % all(exact[k].tobytes() == ref[k].tobytes() for k in ref))
        out.append("ccd 11: " + refusal(m, dict(cols, ccd=np.full(len(cols["t"]), 11, np.int8))))
        out.append("x 2000: " + refusal(m, dict(cols, x=np.full(len(cols["t"]), 2000.0, np.float32))))
        t0 = time.time()
        out.append("70000-event frame: " + refusal(m, {k: np.repeat(v[:1], 70000) for k, v in cols.items()}) + " (%.2f s)" % (time.time() - t0))
        again = m.pileup(cols, alpha, ft, seed)[0]
        out.append("after the refusals identical: %s" % all(again[k].tobytes() == ref[k].tobytes() for k in ref))
    with marx_b200.MarxB200("c3_letg_hrc_s", seed=1, max_photons=1024) as m:
        out.append("HRC context: " + refusal(m, cols))
    open(os.path.join(ROOT, "gpurun_out", "pileup_edges.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
