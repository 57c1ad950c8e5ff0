"""Peak sharding across the GPUs of one box.

Every (mass, threshold) call is independent given the read-only table (reference
mass_explanation.py:116 — the memo is per call), so the path shards with **no data-path collective**:

* the table is replicated (each rank builds it locally from <= 128 weights — 840 B in, 582 MB built on the
  device in well under a millisecond, cheaper than a broadcast);
* peaks are dealt to ranks by mass bin: sort by (window upper bound), deal round-robin.  Enumeration cost
  grows steeply and monotonically with mass, so every rank gets the same share of every cost class;
* each rank runs the ordinary single-GPU batch call on its shard;
* results (status, per-peak counts, fixed-width records) are gathered on the host of rank 0 and put back
  in input order.

``torch.distributed`` is only plumbing here (gloo on CPU in the tests, NCCL's host-object path or gloo on
the GPU box).  The compute callback is injectable so the shard / gather logic is testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def partition_by_mass(masses: np.ndarray, thresholds: Optional[np.ndarray], world: int) -> List[np.ndarray]:
    """Index arrays, one per rank; their union is range(len(masses)), each sorted ascending.

    Peaks are ordered by the upper end of their window (mass + threshold; relative thresholds grow with the
    mass, so mass alone orders them the same way) and dealt round-robin: rank r gets positions r, r+world, ...
    of that order — a mass-bin interleave."""
    masses = np.asarray(masses, dtype=np.float64).reshape(-1)
    key = masses.copy()
    if thresholds is not None:
        t = np.asarray(thresholds, dtype=np.float64).reshape(-1)
        key = key + np.where(np.isnan(t), 0.0, t)
    order = np.argsort(key, kind="stable")
    return [np.sort(order[r::world]) for r in range(world)]


def merge_shards(n_total: int, parts: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]):
    """Inverse of the partition: parts[r] = (indices, status, counts, records[n, W]) of rank r.

    Returns (status[n_total], offsets[n_total+1], records[sum counts, W]) in input order; record widths may
    differ between ranks (each rank sizes W for its own deepest composition) and are 0-padded to the widest."""
    status = np.zeros(n_total, dtype=np.uint8)
    counts = np.zeros(n_total, dtype=np.int64)
    W = max([p[3].shape[1] for p in parts if p[3] is not None and p[3].ndim == 2] + [8])
    for idx, st, cnt, _ in parts:
        status[idx] = st
        counts[idx] = cnt
    offsets = np.zeros(n_total + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    records = np.zeros((int(offsets[-1]), W), dtype=np.uint8)
    for idx, _st, cnt, recs in parts:
        if recs is None or len(recs) == 0:
            continue
        local_off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(cnt, out=local_off[1:])
        # destination row of every local record: offsets[peak] + position inside the peak
        peak_of = np.repeat(np.arange(len(idx)), cnt)
        dst = offsets[idx][peak_of] + (np.arange(len(recs)) - local_off[peak_of])
        records[dst, : recs.shape[1]] = recs
    return status, offsets, records


def explain_masses_sharded(masses, dp_table, max_modifications=np.inf, thresholds=None, with_memo: bool = True,
                           group=None, local_fn: Optional[Callable] = None, dst: int = 0):
    """``explain_masses`` over all ranks of ``group``.  Every rank passes the same full input; rank ``dst``
    gets the merged ``ExplanationBatch`` (input order), the others ``None``.

    ``local_fn(masses, thresholds, max_mods) -> (status, counts, records)`` defaults to the CUDA batch call on
    this rank's device."""
    import torch.distributed as dist

    masses = np.ascontiguousarray(masses, dtype=np.float64).reshape(-1)
    n = len(masses)
    thr = None
    if thresholds is not None:
        thr = (np.full(n, float(thresholds)) if np.ndim(thresholds) == 0
               else np.array([np.nan if x is None else x for x in thresholds], dtype=np.float64))
    mm = None if np.ndim(max_modifications) == 0 else np.asarray(max_modifications)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = partition_by_mass(masses, thr, world)[rank]

    if local_fn is None:
        from . import mass_explanation as ME

        def local_fn(m, t, k):  # noqa: E306
            b = ME.explain_masses(m, dp_table, max_modifications=k, thresholds=t, with_memo=with_memo)
            return b.status, b.counts(), b.records

    st, cnt, recs = local_fn(masses[mine], None if thr is None else thr[mine],
                             max_modifications if mm is None else mm[mine])
    part = (mine, np.asarray(st, dtype=np.uint8), np.asarray(cnt, dtype=np.int64), recs)
    if world == 1:
        parts = [part]
    else:
        parts = [None] * world if rank == dst else None
        dist.gather_object(part, parts, dst=dst, group=group)
        if rank != dst:
            return None
    status, offsets, records = merge_shards(n, parts)
    from .mass_explanation import ExplanationBatch

    weights = np.array([m.mass for m in dp_table.masses], dtype=np.int64)
    return ExplanationBatch(status, offsets, records, weights, [m.names for m in dp_table.masses])


def are_valid_masses_sharded(masses, dp_table, thresholds=None, group=None, local_fn: Optional[Callable] = None,
                             dst: int = 0):
    """``are_valid_masses`` over all ranks: contiguous blocks (validity cost is uniform), gathered on ``dst``."""
    import torch.distributed as dist

    masses = np.ascontiguousarray(masses, dtype=np.float64).reshape(-1)
    n = len(masses)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = [(n * r) // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    thr = None
    if thresholds is not None:
        thr = (np.full(n, float(thresholds)) if np.ndim(thresholds) == 0
               else np.array([np.nan if x is None else x for x in thresholds], dtype=np.float64))
    if local_fn is None:
        from . import mass_explanation as ME

        def local_fn(m, t):  # noqa: E306
            return ME.are_valid_masses(m, dp_table, t)

    mine = np.asarray(local_fn(masses[lo:hi], None if thr is None else thr[lo:hi]), dtype=np.uint8)
    if world == 1:
        return mine
    parts = [None] * world if rank == dst else None
    dist.gather_object(mine, parts, dst=dst, group=group)
    return np.concatenate(parts) if rank == dst else None
