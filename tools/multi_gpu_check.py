#!/usr/bin/env python3
"""Multi-GPU equivalence check (run under torchrun, one rank per GPU): everything goes through the C ABI's own NCCL calls.

G ranks trace the blocks of a few collective steps with marxb200_trace_sharded (time bases all-gathered on the device), the
event lists are merged on rank 0 with marxb200_merge_events_begin / _end (interleaved with the next step's trace, as bench.py
does), tallies are summed with marxb200_tally_allreduce, and rank 0 compares everything with ONE GPU tracing the same rays
step by step: identical events, identical order, identical times.  Steps: three full ones, one with a short last block, two so
small that the upper ranks get empty blocks -- which also walks the time-base look-ahead of marxb200_trace_sharded through
its hits and misses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import marx_b200
from marx_b200 import HISTORY
from marx_b200.api import comm_unique_id, shard_of

MASK = sum(HISTORY[k] for k in ("ENERGY", "TIME", "X_VECTOR", "P_VECTOR", "DET_NUM", "DET_PIXEL", "MIRROR_SHELL", "PULSEHEIGHT",
                                "ORDER", "PI", "SKY_DITHER", "DET_DITHER", "TAG"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")                # only carries the id and the final barrier: the data path is the library's NCCL
    n = 1 << 20                                    # rays per rank and full step
    steps = [n * world, n * world, n * world, n * world - 12345, 70000, 70000]      # look-ahead: miss, hit, hit, miss, miss, hit
    cap = n // 4
    with marx_b200.MarxB200("c2_hetg_acis_s", device=local, seed=77, max_photons=n) as m:
        if os.environ.get("MGC_INIT", "bcast") == "file":
            m.comm_init_file("/tmp/marxb200_comm_%s" % os.environ.get("MASTER_PORT", "0"), rank, world)
        else:
            box = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            m.comm_init(box[0], rank, world)
        specs = [(("order", 23, -11, 12),), (("pha", 1024, 0, 4096),), (("ccd", 10, 0, 10), ("chipx", 64, 0, 1024))]
        tallies = [m.tally_create(*sp) for sp in specs]
        host = np.zeros(cap * world * 96 + 4096, dtype=np.uint8)
        merged, own, first = [], [], 0
        pending = False
        for k, n_total in enumerate(steps):
            f, cnt = m.trace_sharded(first, n_total)
            assert (f, cnt) == shard_of(first, n_total, rank, world)
            if pending:                            # the merge of step k-1 ends while step k is being traced
                lay = m.merge_events_end()
                if rank == 0:
                    cols = m.merge_download(host)
                    merged.append(({c: v.copy() for c, v in cols.items()}, lay))
            for t in tallies:
                t.accumulate()
            own.append(m.download_columns(("time", "ray", "pha")))
            m.merge_events_begin(MASK, 0.0, cap, 0)
            pending = True
            first += n_total
        lay = m.merge_events_end()
        if rank == 0:
            cols = m.merge_download(host)
            merged.append(({c: v.copy() for c, v in cols.items()}, lay))
        for t in tallies:
            t.allreduce()
        summed = [t.read() for t in tallies]
        info = m.comm_info()
        end_time = m.counts()[2]
    # every rank ends with the same running time
    ends = [None] * world
    dist.all_gather_object(ends, end_time)
    assert all(e == ends[0] for e in ends), ends

    if rank == 0:
        with marx_b200.MarxB200("c2_hetg_acis_s", device=local, seed=77, max_photons=n * world) as s:
            first = 0
            all_pha, all_order, all_ccd, all_chipx = [], [], [], []
            for k, n_total in enumerate(steps):
                s.trace(first, n_total)
                ref_abs = s.download_columns(("time", "ray", "pha", "order", "ccd", "chipx"))
                pinned = np.zeros(len(ref_abs["ray"]) * 96 + 4096, dtype=np.uint8)
                s.egress_begin_packed(MASK, 0.0, len(ref_abs["ray"]) + 16)
                ref = s.egress_end_packed(pinned)
                got, lay = merged[k]
                assert lay["n_rows"] == len(ref_abs["ray"]) == sum(lay["rows_of_rank"]), (k, lay["n_rows"], len(ref_abs["ray"]))
                assert set(got) == set(ref)
                for name in ref:
                    if name == "time.dat":
                        # merged TIME = (float) absolute time; the single-GPU file image = (float) ((t - start) + 0) + ... of ITS batch
                        want = ref_abs["time"].astype(np.float32)
                        assert (got[name].astype(np.float32) == want).all(), (k, name)
                    else:
                        assert got[name].tobytes() == ref[name].tobytes(), (k, name)
                # rank 0's own block: the f64 arrival times are bit-identical to the single-GPU trace
                n0 = lay["rows_of_rank"][0]
                assert (own[k]["time"] == ref_abs["time"][:n0]).all() and (own[k]["ray"] == ref_abs["ray"][:n0]).all()
                all_pha.append(ref_abs["pha"]); all_order.append(ref_abs["order"]); all_ccd.append(ref_abs["ccd"]); all_chipx.append(ref_abs["chipx"])
                first += n_total
            assert s.counts()[2] == end_time, (s.counts()[2], end_time)
            o = np.concatenate(all_order).astype(np.int64)
            assert (summed[0] == np.bincount(o + 11, minlength=23)).all()
            assert (summed[1] == np.bincount(np.concatenate(all_pha).astype(np.int64) // 4, minlength=1024)).all()
            img = np.zeros((10, 64), dtype=np.int64)
            np.add.at(img, (np.concatenate(all_ccd).astype(np.int64),
                            np.floor(np.concatenate(all_chipx).astype(np.float64) * (64 / 1024.0)).astype(np.int64)), 1)
            assert (summed[2] == img).all()
        rows = [mm[1]["n_rows"] for mm in merged]
        print("multi_gpu_check OK: world=%d, steps %s -> %s merged events identical to the single-GPU trace (all columns, order, "
              "times); tallies all-reduced in the library equal the binned single-GPU lists; NCCL %d, merge transport: %s; "
              "transfer ms per step %s" % (world, steps, rows, info["nccl_version"], info["merge_transport"],
                                           ["%.3f" % mm[1]["transfer_ms"] for mm in merged]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
