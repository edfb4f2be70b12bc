#!/usr/bin/env python3
"""Multi-GPU equivalence check (run under torchrun, NCCL): G ranks trace G contiguous ray blocks with the
all-gathered time bases, the event lists are merged on rank 0 with gather_event_columns, and rank 0 compares the
result with ONE GPU tracing the same rays in a single batch: identical events, identical order, identical times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import marx_b200
from marx_b200.dist import allreduce_tally, exchange_time_base, gather_event_columns


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 20                                    # per rank; a multiple of 65536 (super-tile aligned)
    names = ("energy", "time", "chipx", "chipy", "pha", "ccd", "order", "ray", "xpos", "ypos", "zpos")
    with marx_b200.MarxB200("c2_hetg_acis_s", device=local, seed=77, max_photons=n * world) as m:
        running = 0.0
        merged_steps = []
        # device-resident tallies, accumulated over the steps and summed over the ranks with one in-place NCCL all-reduce
        specs = [(("order", 23, -11, 12),), (("pha", 1024, 0, 4096),), (("ccd", 10, 0, 10), ("chipx", 64, 0, 1024))]
        tallies = [m.tally_create(*sp) for sp in specs]
        for step in range(2):
            first = (step * world + rank) * n
            base, running = exchange_time_base(m.time_sums(first, n), rank, world, running, device=dev)
            m.create_photons(first, n, base)
            m.mirror_reflect(); m.grating_diffract(); m.detect()
            for t in tallies:
                t.accumulate()
            cols = m.download_columns(names)
            merged = gather_event_columns(cols, rank, world, dst=0, device=dev)
            if rank == 0:
                merged_steps.append(merged)
        # stage-count "histogram" merge: all-reduce over NCCL
        cnt = torch.tensor(m.stage_counts(), device=dev, dtype=torch.int64)
        dist.all_reduce(cnt)
        merged_tallies = [allreduce_tally(t.device_tensor()).cpu().numpy() for t in tallies]
        if rank == 0:
            got = {k: np.concatenate([s[k] for s in merged_steps]) for k in names}
            # the all-reduced tallies equal the binning of the merged event list (= of the single-GPU trace, checked below)
            o = got["order"].astype(np.int64)
            assert (merged_tallies[0] == np.bincount(o + 11, minlength=23)).all()
            assert (merged_tallies[1] == np.bincount(got["pha"].astype(np.int64) // 4, minlength=1024)).all()
            img = np.zeros((10, 64), dtype=np.int64)
            np.add.at(img, (got["ccd"].astype(np.int64), np.floor(got["chipx"].astype(np.float64) * (64 / 1024.0)).astype(np.int64)), 1)
            assert (merged_tallies[2] == img).all()
            assert merged_tallies[0].sum() == len(o)
            m.create_photons(0, n * world, 0.0)
            m.mirror_reflect(); m.grating_diffract(); m.detect()
            a = m.download_columns(names)
            m.trace(n * world, n * world)
            b = m.download_columns(names)
            ref = {k: np.concatenate([a[k], b[k]]) for k in names}
            assert len(got["ray"]) == len(ref["ray"]), (len(got["ray"]), len(ref["ray"]))
            for k in names:
                if k == "time":
                    assert np.abs(got[k] - ref[k]).max() <= 1e-12 * ref[k].max(), k
                else:
                    assert (got[k] == ref[k]).all(), k
            assert (np.diff(got["ray"].astype(np.int64)) > 0).all() and (np.diff(got["time"]) >= 0).all()
            print("multi_gpu_check OK: world=%d, %d events identical to the single-GPU trace; all-reduced tallies (order, PHA, ccd x chipx) "
                  "equal the binned merged list; all-reduced last-step counts %s" % (world, len(got["ray"]), cnt.tolist()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
