"""Host-side statement of the multi-GPU exchanges (torch.distributed; gloo in the CPU tests).  The product path is
marx_b200/csrc/comm.cu behind the C ABI (marxb200_trace_sharded, marxb200_merge_events_begin/_end, marxb200_tally_allreduce: NCCL and
peer writes on device buffers); these functions state the same arithmetic on numpy arrays so that the block rule, the sequential
time-base additions and the rank-order merge can be tested with world_size 2 on a CPU box (tests/test_dist_gloo.py).

The path shards trivially: rank r traces the contiguous block of global ray indices
[(step*G + r)*n, (step*G + r + 1)*n) with the same counter-based draw streams, so every event is
identical for any GPU count.  Two small exchanges remain:

* arrival times are a running sum over ALL rays (source.c:326): ranks all-gather the per-super-tile
  sums of their block (n/65536 doubles) and add the lower ranks' sums sequentially -> `block_time_bases`;
* the per-GPU event lists are merged on demand: blocks are contiguous in ray index and arrival time, so
  the time-ordered merge the reference's marxcat performs (marx/src/marxcat.c:505-535) degenerates to a
  concatenation in rank order -> `gather_event_columns`;
* tallies (device-resident histograms, marxb200_tally_*) are summed over the ranks with one all-reduce on the device
  buffer itself -> `allreduce_tally`.
"""
import numpy as np


def block_time_bases(sums_per_rank, running):
    """sums_per_rank[r] = super-tile sums of rank r's block (canonical order).  Returns (bases, end):
    bases[r] = absolute time at the start of rank r's block, end = time after the last block.  The
    additions are strictly sequential over super-tiles, exactly what a single GPU tracing all blocks does."""
    bases = []
    acc = float(running)
    for sums in sums_per_rank:
        bases.append(acc)
        for v in np.asarray(sums, dtype=np.float64):
            acc = acc + float(v)
    return bases, acc


def exchange_time_base(time_sums, rank, world, running, device=None):
    """all-gather the super-tile sums (tiny) and return (time base of this rank's block, new running time)."""
    import torch
    import torch.distributed as dist
    mine = np.asarray(time_sums, dtype=np.float64).reshape(-1)
    # blocks may be ragged (a short or empty last block): agree on the longest, pad with zeros (adding 0.0 changes no sum), trim again
    cnt = torch.tensor([len(mine)], dtype=torch.int64)
    if device is not None:
        cnt = cnt.to(device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    pad = np.zeros(max(max(counts), 1), dtype=np.float64)
    pad[:len(mine)] = mine
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    sums = [g.cpu().numpy()[:counts[r]] for r, g in enumerate(gathered)]
    bases, end = block_time_bases(sums, running)
    return bases[rank], end


def gather_event_columns(cols, rank, world, dst=0, device=None):
    """cols: dict name -> 1-D numpy array (this rank's events, arrival order).  On `dst` returns the merged
    dict (rank order == arrival-time order); elsewhere None.  Variable lengths are handled by an all-gather of
    the counts followed by padded gathers."""
    import torch
    import torch.distributed as dist
    names = sorted(cols)
    n_local = len(cols[names[0]]) if names else 0
    cnt = torch.tensor([n_local], dtype=torch.int64)
    if device is not None:
        cnt = cnt.to(device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    out = {} if rank == dst else None
    for name in names:
        a = np.ascontiguousarray(cols[name])
        raw = a.view(np.uint8).reshape(n_local, a.dtype.itemsize) if n_local else np.zeros((0, a.dtype.itemsize), np.uint8)
        pad = np.zeros((nmax, a.dtype.itemsize), dtype=np.uint8)
        pad[:n_local] = raw
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        if rank == dst:
            parts = [bufs[r][:counts[r]].cpu().numpy().reshape(-1).view(a.dtype) for r in range(world)]
            out[name] = np.concatenate(parts)
    return out


def allreduce_tally(counts):
    """counts: an integer torch tensor -- on the GPU box the alias of a tally's device buffer (Tally.device_tensor(),
    NCCL sums in place over NVLink, no host copy), in the CPU tests a gloo tensor.  After the call every rank holds the
    sum over all ranks: integer counts, so the merged histogram is exactly what one GPU tracing all rays accumulates."""
    import torch.distributed as dist
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
