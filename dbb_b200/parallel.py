"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL over NVLink).

Stage 1 (counting) shards by independent scaffolds: longest-processing-time-first over total base pairs,
the same heuristic the reference uses for its worker processes (preprocess.py:373-387).  No collective.
Hand-off: ONE all-gather of the float32 signature matrix plus one of the per-scaffold scalars (length,
GC, coverage, original id: ``all_gather_scalars``), so a rank only ever needs the metadata of its own shard.  Stage 2 (pairs / kNN) shards by equal row blocks over the gathered order;
every rank reads all columns; per-row results are disjoint, so there is no reduction afterwards.

All partition logic is pure host code (tested on CPU with gloo, world_size 2).
"""
import heapq

import numpy as np


def lpt_partition(lengths, world):
    """Assign scaffolds to ``world`` ranks: descending length, each to the rank with the fewest total bp
    (preprocess.py:380-387).  Returns a list of index arrays (each ascending) -- exact LPT up to 2M items,
    a snake deal of the sorted order beyond that (indistinguishable for millions of short contigs)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = lengths.shape[0]
    order = np.argsort(-lengths, kind='stable')
    if world == 1:
        return [np.arange(n, dtype=np.int64)]
    owner = np.empty(n, dtype=np.int32)
    if n <= 2000000:
        heap = [(0, r) for r in range(world)]
        for idx in order:
            tot, r = heapq.heappop(heap)
            owner[idx] = r
            heapq.heappush(heap, (tot + int(lengths[idx]), r))
    else:
        pos = np.arange(n) % (2 * world)
        owner[order] = np.where(pos < world, pos, 2 * world - 1 - pos)
    return [np.nonzero(owner == r)[0].astype(np.int64) for r in range(world)]


def row_block(n, rank, world):
    """Equal row blocks [r0, r1) of the gathered order (cost per row is uniform: n columns)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gathered_order(parts):
    """Scaffold ids in gathered order (rank-major, local order inside a rank) and its inverse."""
    order = np.concatenate(parts) if len(parts) else np.zeros(0, dtype=np.int64)
    inv = np.empty_like(order)
    inv[order] = np.arange(order.shape[0])
    return order, inv


def imbalance(lengths, parts):
    """max / mean of per-rank base pairs."""
    tot = np.array([int(np.asarray(lengths)[p].sum()) for p in parts], dtype=np.float64)
    return float(tot.max() / tot.mean()) if tot.mean() > 0 else 1.0


def all_gather_rows(local, counts, group=None):
    """All-gather of row-sharded 2-D tensors with uneven row counts (``counts[r]`` rows on rank r) into
    one [sum(counts), width] tensor in gathered order.  NCCL on GPUs, gloo on CPU tensors."""
    import torch
    import torch.distributed as dist
    world = len(counts)
    if world == 1:
        return local
    n_max = int(max(counts))
    width = local.shape[1]
    pad = local
    if local.shape[0] != n_max:
        pad = torch.zeros((n_max, width), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    buf = torch.empty((world, n_max, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(world * n_max, width), pad.contiguous(), group=group)
    if all(int(c) == n_max for c in counts):
        return buf.view(world * n_max, width)
    return torch.cat([buf[r, :int(counts[r])] for r in range(world)], dim=0)


def all_gather_scalars(ids, length, gc=None, coverage=None, counts=None, device=None, group=None):
    """The K1 -> K2 hand-off of the per-scaffold scalars: every rank contributes the original ids,
    lengths and (optionally) GC / coverage of ITS shard only; all ranks get the arrays of all scaffolds in gathered
    order (rank-major, the order of ``all_gather_rows``).  One collective over a packed float64 [n, 4] record
    (ids and lengths are exact in float64 up to 2^53).  ``counts[r]`` = scaffolds of rank r (every rank knows them: they
    are the row counts of the signature gather).  Returns (ids int64, length int64, gc float64 | None, coverage | None)."""
    import torch
    ids = np.asarray(ids, dtype=np.int64)
    rec = np.zeros((ids.shape[0], 4), dtype=np.float64)
    rec[:, 0] = ids
    rec[:, 1] = np.asarray(length, dtype=np.float64)
    if gc is not None:
        rec[:, 2] = gc
    if coverage is not None:
        rec[:, 3] = coverage
    t = torch.from_numpy(rec)
    if device is not None:
        t = t.to(device)
    if counts is None:
        counts = [ids.shape[0]]
    full = all_gather_rows(t, counts, group=group).cpu().numpy()
    return (full[:, 0].astype(np.int64), full[:, 1].astype(np.int64), full[:, 2].copy() if gc is not None else None,
            full[:, 3].copy() if coverage is not None else None)
