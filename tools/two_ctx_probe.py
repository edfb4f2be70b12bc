#!/usr/bin/env python3
"""How much of a traced batch's time is kernel-boundary idle?  One context tracing K batches back to back, against TWO contexts
(own streams, own buffers) tracing alternate batches of the same size: the second stream's kernels fill the SM slots the first
one's persistent kernels leave idle while they drain.  Prints rays/s of both arrangements."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import marx_b200


def run(ms, n, steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        ms[s % len(ms)].trace(s * n, n)
    torch.cuda.synchronize()
    return steps * n / (time.perf_counter() - t0)


def main():
    n = 1 << 24
    steps = int(os.environ.get("STEPS", "40"))
    torch.cuda.set_device(0)
    a = marx_b200.MarxB200("c2_hetg_acis_s", seed=1, max_photons=n)
    b = marx_b200.MarxB200("c2_hetg_acis_s", seed=1, max_photons=n)
    for m in (a, b):
        for s in range(3):
            m.trace(s * n, n)
    one = max(run([a], n, steps) for _ in range(3))
    two = max(run([a, b], n, steps) for _ in range(3))
    print("one context %.4g rays/s, two contexts alternating %.4g rays/s (%+.1f %%)" % (one, two, 100 * (two / one - 1)))
    a.close(); b.close()


if __name__ == "__main__":
    main()
