"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: canonical time bases and the event merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marx_b200.dist import block_time_bases, exchange_time_base, gather_event_columns


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    # this rank's super-tile sums; RAGGED: the last block of a step may be short (fewer super-tiles) or empty
    n_sums = [7, 4, 0][rank] if world == 3 else [7, 3][rank]
    sums = rng.random(n_sums) * 3.0 + rank
    base, end = exchange_time_base(sums, rank, world, running=10.0)
    # events of this rank: times inside its block, ragged counts (rank 1 has none -> empty-input edge case)
    n_ev = [5, 0, 3][rank] if world == 3 else [5, 0][rank]
    t = np.sort(base + rng.random(n_ev) * (sums.sum() if n_sums else 1.0))
    cols = {"time": t, "pha": (np.arange(n_ev) + 100 * rank).astype(np.int16), "ray": (np.arange(n_ev) + 1000 * rank).astype(np.uint64)}
    merged = gather_event_columns(cols, rank, world, dst=0)
    q.put((rank, sums, base, end, cols, merged))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_time_base_exchange_and_event_merge(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sums = [r[1] for r in res]
    bases, end = block_time_bases(sums, 10.0)
    # sequential (single-GPU) order of additions
    acc = 10.0
    for r in range(world):
        assert res[r][2] == acc == bases[r]
        for v in sums[r]:
            acc += v
        assert res[r][3] == end
    assert acc == end
    merged = res[0][5]
    for name in ("time", "pha", "ray"):
        want = np.concatenate([res[r][4][name] for r in range(world)])
        assert merged[name].dtype == want.dtype and (merged[name] == want).all()
    assert (np.diff(merged["time"]) >= 0).all()          # rank order == arrival order
    assert all(res[r][5] is None for r in range(1, world))


def _tally_worker(rank, world, port, q):
    from marx_b200.dist import allreduce_tally
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7 + rank)
    local = rng.integers(0, 1000, size=(10, 7)).astype(np.int64)
    if rank == 1:
        local[:] = 0                             # a rank without events
    merged = allreduce_tally(torch.from_numpy(local.copy()))
    q.put((rank, local, merged.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tally_allreduce(world):
    """per-rank histograms are summed in place on every rank (the GPU path hands the device buffer itself to NCCL)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tally_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(r[1] for r in res)
    for r in res:
        assert (r[2] == total).all()
