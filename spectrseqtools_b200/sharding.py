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


# ----------------------------------------------------------------------------- contiguous blocks + shared-memory gather
def estimated_compositions(masses: np.ndarray, thresholds: Optional[np.ndarray], dp_table) -> np.ndarray:
    """Host-side cost of every call for the partition: window size times the density of compositions around the
    mass.  The density comes from a coarse coin-change count over the row weights (bucket = a fifth of the lightest
    row) — no device work, no table access; it only has to rank the calls (a 5-nt difference costs 10^4 times a
    1-nt one)."""
    masses = np.asarray(masses, dtype=np.float64).reshape(-1)
    w = np.array([m.mass for m in dp_table.masses if m.mass > 0], dtype=np.int64)
    if len(w) == 0 or len(masses) == 0:
        return np.ones(len(masses))
    finite = masses[np.isfinite(masses)]
    top = int(min(max(float(finite.max()) if len(finite) else 0.0, 0.0) / dp_table.precision, 1e15)) + int(w.max())
    width = max(1, int(w.min()) // 5, top // 4096 + 1)
    K = top // width + 2
    dens = np.zeros(K)
    dens[0] = 1.0
    buckets = np.maximum(w // width, 1)
    for x in np.unique(buckets):  # unbounded coin change over the bucketed weights (rows sharing a bucket: multiplicity)
        mult, step = float(np.sum(buckets == x)), int(x)
        for start in range(step, K, step):
            dens[start:start + step] += mult * dens[start - step:start][: K - start]
    thr = dp_table.tolerance * masses if thresholds is None else np.where(np.isnan(thresholds), dp_table.tolerance * masses, thresholds)
    win = 2.0 * np.ceil(thr / dp_table.precision) + 1.0
    k = np.clip(np.nan_to_num(masses / dp_table.precision / width, nan=0.0, posinf=K - 1, neginf=0.0), 0, K - 1).astype(np.int64)
    return 16.0 + win * (1.0 + dens[k] / width)


def partition_contiguous(masses, thresholds, world: int, dp_table, counts: Optional[np.ndarray] = None) -> List[int]:
    """Cut points [0 = c_0 <= c_1 <= ... <= c_world = n]: rank r takes calls c_r .. c_{r+1}.  Blocks are contiguous in
    INPUT order (so gathering is a concatenation: no per-record scatter on the host) and equal in work.

    ``counts``: compositions per call looked up on the device (``mass_explanation.count_compositions``: exact when no
    budget binds) — the work of a call is then a constant plus its compositions; calls whose count is not known
    (2**64 - 1) and a missing ``counts`` fall back to the host-side estimate."""
    n = len(masses)
    if n == 0:
        return [0] * (world + 1)
    t = None if thresholds is None else np.asarray(thresholds, dtype=np.float64).reshape(-1)
    work = estimated_compositions(masses, t, dp_table)
    if counts is not None:
        c = np.asarray(counts, dtype=np.uint64)
        known = c != np.uint64(2**64 - 1)
        work = np.where(known, 16.0 + c.astype(np.float64), work)
    cost = np.cumsum(work)
    cuts = [0] + [int(np.searchsorted(cost, cost[-1] * r / world)) for r in range(1, world)] + [n]
    return [int(x) for x in np.maximum.accumulate(np.minimum(cuts, n))]


class ShmGather:
    """Host-side gather on ONE box through POSIX shared memory: every rank owns a segment, puts its arrays plus a small
    header into it, and rank 0 maps all segments and reads them in place (views, no pickling, no socket).

    The segment is cut into ``REGIONS`` regions used round-robin by consecutive steps (numbered from 1).  A rank can
    have its results written STRAIGHT INTO a region — ``register(ctx)`` page-locks the segment, ``region(seq)`` is the
    buffer to hand to ``explain_masses(out_block=...)`` / ``classify_observed(out=...)`` — so that the device-to-host
    copy is the gather and ``publish`` only writes a header; arrays that live elsewhere are copied in.
    ``publish(seq, arrays)`` then ``collect(seq)`` on rank 0.  Rank 0 hands out VIEWS: a region is written again only
    after rank 0 has released the step that used it (``collect(seq + 1)`` releases ``seq``)."""

    HEADER = 8192
    REGIONS = 8  # a rank with d steps in flight needs d + 2 (the step it submits, the d - 1 behind it, the one rank 0 still reads)
    SLOT = 96  # header words per region: count + 5 per array (at most 19 arrays)

    def __init__(self, tag: str, rank: int, world: int, capacity: int = 256 << 20):
        from multiprocessing import shared_memory

        self.rank, self.world, self.tag = rank, world, tag
        self._rs = (capacity // self.REGIONS) & ~4095
        self._shm = shared_memory.SharedMemory(name=f"{tag}_{rank}", create=True, size=self.HEADER + self._rs * self.REGIONS)
        self._hdr = np.ndarray(self.HEADER // 8, dtype=np.int64, buffer=self._shm.buf)
        self._hdr[:] = 0
        self._data = np.ndarray(self._rs * self.REGIONS, dtype=np.uint8, buffer=self._shm.buf, offset=self.HEADER)
        self._base = self._data.__array_interface__["data"][0]
        self._peers = {}
        self._sm = shared_memory
        self._registered = None

    def register(self, ctx) -> None:
        """Page-lock this rank's segment (``_cabi.Context.host_register``) so that results can be copied straight into it."""
        ctx.host_register(self._data)
        self._registered = ctx

    def _peer(self, r: int):
        if r == self.rank:
            return self._hdr, self._data
        if r not in self._peers:
            import time

            deadline = time.time() + 60.0
            while True:
                try:
                    shm = self._sm.SharedMemory(name=f"{self.tag}_{r}")
                    break
                except FileNotFoundError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.001)
            hdr = np.ndarray(self.HEADER // 8, dtype=np.int64, buffer=shm.buf)
            data = np.ndarray(shm.size - self.HEADER, dtype=np.uint8, buffer=shm.buf, offset=self.HEADER)
            self._peers[r] = (shm, hdr, data)
        _shm, hdr, data = self._peers[r]
        return hdr, data

    @staticmethod
    def _spin(waiting, what: str, timeout: float = 60.0) -> None:
        """Busy-wait while ``waiting()``; a peer that never arrives is an error, not a hang."""
        import time

        n, deadline = 0, None
        while waiting():
            n += 1
            if n & 0xFFFF == 0:
                now = time.monotonic()
                deadline = deadline or now + timeout
                if now > deadline:
                    raise TimeoutError(f"ShmGather: {what} within {timeout:.0f} s")

    def region(self, seq: int) -> np.ndarray:
        """The region step ``seq`` may write (uint8 view), once rank 0 has released the step that used it last."""
        hdr = self._hdr
        self._spin(lambda: seq > self.REGIONS and hdr[1] < seq - self.REGIONS, f"rank 0 never released step {seq - self.REGIONS}")
        k = seq % self.REGIONS
        return self._data[k * self._rs:(k + 1) * self._rs]

    def publish(self, seq: int, arrays: Sequence[np.ndarray]) -> None:
        """header: [0] seq (written last), [1] released seq (written by rank 0), then per region a slot of
        [number of arrays, (dtype code, ndim, shape0, shape1, byte offset) per array]."""
        hdr = self._hdr
        reg = self.region(seq)
        r0 = (seq % self.REGIONS) * self._rs
        if len(arrays) * 5 + 1 > self.SLOT:
            raise ValueError("too many arrays for one step")
        # arrays that already live in this step's region stay where they are; the others are copied behind them
        placed, cursor = [], 0
        for a in arrays:
            a = np.ascontiguousarray(a)
            off = a.__array_interface__["data"][0] - self._base - r0 if a.size else -1
            inside = 0 <= off and off + a.nbytes <= self._rs
            placed.append((a, off if inside else None))
            if inside:
                cursor = max(cursor, off + a.nbytes)
        slot = 8 + (seq % self.REGIONS) * self.SLOT
        hdr[slot] = len(arrays)
        for k, (a, off) in enumerate(placed):
            if off is None:
                cursor = (cursor + 63) & ~63
                if cursor + a.nbytes > self._rs:
                    raise MemoryError("ShmGather region too small for this step")
                reg[cursor:cursor + a.nbytes] = a.reshape(-1).view(np.uint8)
                off = cursor
                cursor += a.nbytes
            hdr[slot + 1 + 5 * k: slot + 6 + 5 * k] = [_DT.index(a.dtype.str), a.ndim, a.shape[0] if a.ndim else 1,
                                                      a.shape[1] if a.ndim > 1 else 1, r0 + off]
        hdr[0] = seq

    def release(self, seq: int) -> None:
        """Rank 0 is done with the views of every step up to ``seq``: their regions may be written again."""
        for r in range(self.world):
            hdr = self._peer(r)[0]
            if hdr[1] < seq:
                hdr[1] = seq

    def collect(self, seq: int):
        """Rank 0: the arrays of every rank for step ``seq`` (views of the shared segments), in rank order.  The views
        stay valid until the next ``collect`` (which releases step ``seq - 1``) or an explicit ``release``."""
        self.release(seq - 1)
        out = []
        for r in range(self.world):
            hdr, data = self._peer(r)
            self._spin(lambda: hdr[0] < seq, f"rank {r} never published step {seq}")
            slot = 8 + (seq % self.REGIONS) * self.SLOT
            arrs = []
            for k in range(int(hdr[slot])):
                code, ndim, s0, s1, at = (int(x) for x in hdr[slot + 1 + 5 * k: slot + 6 + 5 * k])
                dt = np.dtype(_DT[code])
                n = s0 * (s1 if ndim > 1 else 1)
                a = data[at:at + n * dt.itemsize].view(dt)
                arrs.append(a.reshape(s0, s1) if ndim > 1 else a)
            out.append(arrs)
        return out

    def close(self) -> None:
        if self._registered is not None:
            try:
                self._registered.host_unregister(self._data)
            except Exception:
                pass
            self._registered = None
        for shm, _h, _d in self._peers.values():
            shm.close()
        self._peers = {}
        self._hdr = self._data = None
        try:
            self._shm.close()
            self._shm.unlink()
        except Exception:
            pass


_DT = ["|u1", "<u4", "<i8", "<u8", "<f8", "<i4"]
