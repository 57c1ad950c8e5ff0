"""Mass explanation on the B200 behind the reference's ``spectrseqtools.mass_explanation`` names.

Scalar entry points keep the reference's signatures and result structures
(``is_valid_mass`` mass_explanation.py:45-89, ``explain_mass_with_table`` :92-203,
``explain_mass_with_recursion`` :206-284, ``convert_nucleotide_masses_to_names`` :287-320,
``MassExplanations`` :12-14, ``MASS_NAMES`` :17-27, ``IS_MOD`` :29-42) so that prediction.py,
skeleton_building.py and linear_program.py can consume them unchanged.  ``explain_masses`` and
``are_valid_masses`` are the batched forms the GPU is built for: many (mass, threshold) pairs per call,
results as CSR arrays, Python sets only on demand.

Host work here is limited to what must match CPython bit for bit (float -> integer conversions, banker's
rounding of budgets) and to shaping results.  All table reads, window scans and enumeration run in
libsst_b200.so; without it (or without a B200) every call raises ``_cabi.DeviceUnavailable``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from itertools import chain, combinations_with_replacement, product
from typing import Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from . import _cabi
from .mass_table import DynamicProgrammingTable
from .masses import _INT_MASS_IS_MOD, _INT_MASS_NAMES
from .mass_table import NucleotideMass as _NucleotideMass, row_version as _row_version


@dataclass
class MassExplanations:
    explanations: Set[Tuple[str]]


# integer mass -> representative nucleosides / modification flag, for the GLOBAL alphabet
MASS_NAMES = {m: list(names) for m, names in _INT_MASS_NAMES.items()}
IS_MOD = dict(_INT_MASS_IS_MOD)


# ---------------------------------------------------------------- host-side integer conversions

def _integerise(mass: float, threshold: Optional[float], dp_table) -> Tuple[int, int]:
    """(target, threshold) in table units, the way mass_explanation.py:51-58 / :107-114 compute them."""
    target = int(round(mass / dp_table.precision, 0))
    if threshold is None:
        threshold = dp_table.tolerance * mass
    return target, int(np.ceil(threshold / dp_table.precision))


def _integerise_many(masses: np.ndarray, thresholds, dp_table) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised ``_integerise``: IEEE division + round-half-even (np.rint == CPython round(x, 0)) + ceil,
    element for element the same float operations as the scalar code (checked in tests/test_host.py)."""
    masses = np.asarray(masses, dtype=np.float64)
    precision = float(dp_table.precision)
    target = np.rint(masses / precision).astype(np.int64)
    rel = float(dp_table.tolerance) * masses
    if thresholds is None:
        thr_f = rel
    else:
        if np.ndim(thresholds) == 0:
            thresholds = [thresholds] * len(masses)
        t = np.array([np.nan if x is None else x for x in thresholds], dtype=np.float64) if not isinstance(thresholds, np.ndarray) else thresholds.astype(np.float64)
        thr_f = np.where(np.isnan(t), rel, t)
    thr = np.ceil(thr_f / precision).astype(np.int64)
    return target, thr


def _budget_int(x) -> int:
    """Budgets are only tested with ``> 0`` and decremented by one: ceil() keeps that behaviour for floats."""
    if x is None:
        return _cabi.BUDGET_INF
    if isinstance(x, float) and math.isnan(x):
        return 0  # upstream tests the budget with `> 0`: NaN never allows a modification
    if isinstance(x, float) and math.isinf(x):
        return _cabi.BUDGET_INF if x > 0 else 0
    v = math.ceil(x)
    return max(0, min(int(v), _cabi.BUDGET_INF))


def _row_metadata(dp_table) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """weights, is_mod, IND[r] = round(max_len * rate_r) (Python round, mass_explanation.py:158-161,200).

    Cached on the table object; the key holds everything the arrays depend on (callers mutate rates and
    ``seq.max_len`` between calls, and alphabet reduction swaps the row list)."""
    rows = dp_table.masses
    if rows and type(rows[0]) is _NucleotideMass:
        key = (id(rows), len(rows), dp_table.seq.max_len, _row_version())  # (any rate assignment anywhere bumps the version)
    else:  # duck-typed rows: read the rates
        key = (id(rows), len(rows), dp_table.seq.max_len, tuple(m.modification_rate for m in rows))
    hit = getattr(dp_table, "_row_meta", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    weights = np.array([m.mass for m in rows], dtype=np.int64)
    is_mod = np.array([1 if m.is_modification else 0 for m in rows], dtype=np.uint8)
    ind = np.array([_budget_int(round(dp_table.seq.max_len * m.modification_rate)) for m in rows], dtype=np.int32)
    try:
        dp_table._row_meta = (key, (weights, is_mod, ind))
    except AttributeError:  # duck-typed tables without a __dict__
        pass
    return weights, is_mod, ind


def _modes(weights, is_mod, ind, max_mods: np.ndarray, hi: np.ndarray, with_memo: bool) -> np.ndarray:
    """Per-peak budget mode.  FREE when no composition inside the window can exhaust a budget:
    max_mods >= hi // min modified weight, and IND[r] >= hi // w_r for every modified row r, i.e.
    hi < min_r (IND[r] + 1) * w_r.  (Same rule as sst_explain_stage_f64.)"""
    mod_rows = np.nonzero(np.asarray(is_mod)[1:])[0] + 1
    slow = _cabi.MODE_MEMO if with_memo else _cabi.MODE_EXACT
    if len(mod_rows) == 0:
        return np.zeros(len(hi), dtype=np.uint8)
    w_mod = np.asarray(weights, dtype=np.int64)[mod_rows]
    hi_pos = np.maximum(np.asarray(hi, dtype=np.int64), 0)
    hi_limit = int(((np.asarray(ind, dtype=np.int64)[mod_rows] + 1) * w_mod).min())
    free = (np.asarray(max_mods, dtype=np.int64) >= hi_pos // int(w_mod.min())) & (hi_pos < hi_limit)
    return np.where(free, _cabi.MODE_FREE, slow).astype(np.uint8)


# ---------------------------------------------------------------- batched API

class ExplanationBatch:
    """Result of ``explain_masses``: compositions of all peaks as fixed-width row-index records.

    ``offsets[p] : offsets[p+1]`` are the records of peak p; each record holds table-row indices in
    ascending order, 0-padded.  ``status`` carries the per-peak flags (zero in window / out of table).
    """

    def __init__(self, status, offsets, records, weights, names_by_row):
        self.status = status
        self._records = records  # uint8[n, W], or the planes a queued batch crosses the bus as (_cabi.SplitRecords)
        self._offsets = offsets  # int64, or the uint32 the asynchronous entry brings back (widened on first use)
        self.weights = weights
        self._names_by_row = names_by_row

    @property
    def records(self) -> np.ndarray:
        """uint8[n, W]: row indices ascending, 0-padded (planes are put together on first use)."""
        if isinstance(self._records, _cabi.SplitRecords):
            self._records = self._records.materialize()
        return self._records

    def raw_records(self):
        """The arrays the records arrived in: [uint8[n, W]], or [lo uint32[n], plane uint8[n], ...] of a queued batch."""
        r = self._records
        return [r.lo] + r.planes if isinstance(r, _cabi.SplitRecords) else [r]

    @property
    def offsets(self) -> np.ndarray:
        if self._offsets.dtype != np.int64:
            self._offsets = self._offsets.astype(np.int64)
        return self._offsets

    def __len__(self):
        return len(self.status)

    @property
    def n_compositions(self) -> int:
        return int(self._offsets[-1])

    def counts(self) -> np.ndarray:
        return np.diff(self.offsets)

    def out_of_table(self, p: int) -> bool:
        return bool(self.status[p] & _cabi.STATUS_OUT_OF_TABLE)

    def has_solution(self, p: int) -> bool:
        """False = the reference would return ``MassExplanations(None)``."""
        return bool(self.status[p] & _cabi.STATUS_ZERO_IN_WINDOW) or self.offsets[p + 1] > self.offsets[p]

    def rows(self, p: int) -> np.ndarray:
        return self.records[self.offsets[p]:self.offsets[p + 1]]

    def solutions(self, p: int) -> List[List[int]]:
        """Integer-weight lists in ascending order, like the reference's ``solutions`` (:191-201)."""
        recs = self.rows(p)
        w = self.weights
        out = [[int(w[r]) for r in rec if r] for rec in recs]
        if self.status[p] & _cabi.STATUS_ZERO_IN_WINDOW:
            out.append([])
        return out

    def explanations(self, p: int) -> MassExplanations:
        if self.out_of_table(p):
            raise _cabi.TableTooSmall("A value of the mass window is not in the DP table. Extend its size if you want to compute larger masses.")
        return convert_nucleotide_masses_to_names(self.solutions(p))

    def canonical(self, p: int) -> List[Tuple[int, ...]]:
        """Sorted tuples of row indices (device-side canonical form for parity checks)."""
        return sorted(tuple(int(r) for r in rec if r) for rec in self.rows(p))


def _thr_array(thresholds, n: int) -> Optional[np.ndarray]:
    """Per-mass absolute thresholds as float64, NaN where the reference would use ``tolerance * mass``."""
    if thresholds is None:
        return None
    if isinstance(thresholds, np.ndarray) and thresholds.dtype == np.float64:
        return thresholds.reshape(-1)
    if np.ndim(thresholds) == 0:
        return np.full(n, float(thresholds), dtype=np.float64)
    return np.array([np.nan if x is None else x for x in thresholds], dtype=np.float64)


class PendingExplanations:
    """A batch submitted with ``explain_masses(..., wait=False)``: ``wait()`` gives the ``ExplanationBatch``."""

    def __init__(self, ctx, dp_table, weights, copy):
        self._ctx, self._dp, self._weights, self._copy = ctx, dp_table, weights, copy
        self._batch = None

    def wait(self) -> "ExplanationBatch":
        if self._batch is None:
            status, off, recs = self._ctx.explain_collect()
            if self._copy:
                status, off = status.copy(), off.astype(np.int64)
                recs = recs.materialize() if isinstance(recs, _cabi.SplitRecords) else recs.copy()
            self._batch = ExplanationBatch(status, off, recs, self._weights, [m.names for m in self._dp.masses])
        return self._batch


def explain_masses(masses: Sequence[float], dp_table: DynamicProgrammingTable, max_modifications=np.inf,
                   thresholds=None, with_memo: bool = True, compression_rate: Optional[int] = None,
                   fetch_records: bool = True, copy: bool = True, wait: bool = True, slot: int = 0, out_block=None):
    """Batched ``explain_mass_with_table``: one device pass for all masses.

    ``max_modifications`` and ``thresholds`` may be scalars or per-mass sequences (``thresholds`` None, or a
    None / NaN entry, = relative ``dp_table.tolerance * mass``).  The float -> integer conversion and the
    choice of the budget mode happen inside the library (``sst_explain_stage_f64``) with the reference's
    float operations.  ``copy=False`` returns views of the context's pinned result buffers, valid until
    the next call on the same device and slot.

    ``wait=False`` (one modification budget for the batch) queues the whole call — inputs in, staging, pass, results
    out — and returns a ``PendingExplanations`` at once; with a different ``slot`` per call two batches are in flight
    on the device, the copies of one under the kernels of the other (``masses`` / ``thresholds`` must not be modified
    until ``wait()`` returns).  ``out_block``: a page-locked uint8 buffer of the caller's (``_cabi.Context.host_register``)
    the result block is copied into instead of the context's own — e.g. a region of a shared-memory segment, so that the
    other processes of the box see the result without a further copy (``sharding.ShmGather``).
    """
    if compression_rate is not None and compression_rate != dp_table.compression_per_cell:
        raise ValueError("compression_rate must match the table's compression_per_cell")
    masses = np.ascontiguousarray(masses, dtype=np.float64).reshape(-1)
    P = len(masses)
    thr = _thr_array(thresholds, P)
    if np.ndim(max_modifications) == 0:
        max_mods = _budget_int(max_modifications)  # one budget for the batch: filled on the device
    else:
        max_mods = np.array([_budget_int(x) for x in max_modifications], dtype=np.int32)
    dev = dp_table.device_table()
    ctx = dev.ctx if slot == 0 else _cabi.context(dev.ctx.device, slot)
    weights, is_mod, ind = _row_metadata(dp_table)
    if not wait:
        if np.ndim(max_modifications) != 0:
            raise ValueError("wait=False takes one modification budget for the whole batch")
        ctx.explain_submit_f64(dev, masses, thr, max_mods, ind, is_mod, dp_table.precision, dp_table.tolerance, with_memo, out_block=out_block)
        return PendingExplanations(ctx, dp_table, weights, copy)
    ctx.explain_stage_f64(dev, masses, thr, max_mods, ind, is_mod, dp_table.precision, dp_table.tolerance, with_memo)
    return _run_and_fetch(dp_table, dev, weights, fetch_records, copy, ctx)


def count_compositions(masses: Sequence[float], dp_table: DynamicProgrammingTable, thresholds=None) -> np.ndarray:
    """How many compositions ``explain_mass_with_table`` would return per mass when no modification budget binds — looked
    up in the alphabet's composition-count table, nothing is enumerated.  uint64; ``2**64 - 1`` where the answer is not
    known (window beyond the count table, saturated count, non-finite input).  What ``sharding.partition_contiguous``
    balances the GPUs of a box by."""
    masses = np.ascontiguousarray(masses, dtype=np.float64).reshape(-1)
    dev = dp_table.device_table()
    return dev.ctx.count_compositions_f64(dev, masses, _thr_array(thresholds, len(masses)), dp_table.precision, dp_table.tolerance)


def _run_and_fetch(dp_table, dev, weights, fetch_records=True, copy=True, ctx=None) -> ExplanationBatch:
    ctx = ctx or dev.ctx
    cap = 0
    while True:
        try:
            ctx.explain_run(dev, 0, cap)
            break
        except _cabi.MemoFull:
            cap = (cap or (1 << 20)) * 4
            if cap > (1 << 30):
                raise
    status, off, recs = ctx.explain_fetch(want_records=fetch_records, copy=copy)
    return ExplanationBatch(status, off, recs, weights, [m.names for m in dp_table.masses])


def _explain_integer(dp_table, target, thr, max_mods, with_memo=True, fetch_records=True) -> ExplanationBatch:
    """Integer-domain entry (targets / thresholds already in table units); budget modes chosen here."""
    dev = dp_table.device_table()
    weights, is_mod, ind = _row_metadata(dp_table)
    target = np.asarray(target, dtype=np.int64)
    thr = np.asarray(thr, dtype=np.int64)
    max_mods = np.asarray(max_mods, dtype=np.int32)
    mode = _modes(weights, is_mod, ind, max_mods.astype(np.int64), target + thr, with_memo)
    dev.ctx.explain_stage(dev, target, thr, max_mods, mode, ind, is_mod)
    return _run_and_fetch(dp_table, dev, weights, fetch_records)


def are_valid_masses(masses: Sequence[float], dp_table: DynamicProgrammingTable, thresholds=None) -> np.ndarray:
    """Batched ``is_valid_mass`` -> uint8 array of _cabi.VALID_* codes (2 = out of table)."""
    masses = np.ascontiguousarray(masses, dtype=np.float64).reshape(-1)
    thr = _thr_array(thresholds, len(masses))
    dev = dp_table.device_table()
    ctx = dev.ctx
    if len(masses) <= 4096:  # small batches (the reference-shaped call is a batch of one): one call, one synchronisation
        return ctx.is_valid_f64(dev, masses, thr, dp_table.precision, dp_table.tolerance)
    ctx.valid_stage_f64(masses, thr, dp_table.precision, dp_table.tolerance)
    ctx.valid_run(dev)
    return ctx.valid_fetch()


# ---------------------------------------------------------------- reference-shaped scalar API

def is_valid_mass(mass: float, dp_table: DynamicProgrammingTable, threshold: float = None) -> bool:
    code = int(are_valid_masses([mass], dp_table, None if threshold is None else [threshold])[0])
    if code == _cabi.VALID_OUT_OF_TABLE:
        raise _cabi.TableTooSmall("A value of the mass window is not in the DP table. Extend its size if you want to compute larger masses.")
    return code == _cabi.VALID_YES


def explain_mass_with_table(mass: float, dp_table: DynamicProgrammingTable, max_modifications=np.inf,
                            compression_rate=None, threshold=None, with_memo=True) -> MassExplanations:
    """Return all possible combinations of nucleosides that could sum up to the given mass."""
    # a batch of one through the queued entry: inputs in, staging, pass and the result block out with ONE wait (the
    # staged path waits three times; a call whose budget can bind is redone inside wait(), as for any batch)
    batch = explain_masses([mass], dp_table, max_modifications=max_modifications,
                           thresholds=None if threshold is None else [threshold], with_memo=with_memo,
                           compression_rate=compression_rate, wait=False).wait()
    return batch.explanations(0)


def convert_nucleotide_masses_to_names(solutions: List[List[int]]) -> MassExplanations:
    """Integer-weight lists -> set of name tuples (reference :287-320).

    ``None`` when there is no solution at all; the empty solution contributes nothing; a weight shared by
    several representatives expands into every multiset of their names.
    """
    if len(solutions) == 0:
        return MassExplanations(None)
    named = set()
    for sol in solutions:
        if len(sol) == 0:
            continue
        # one pool per run of equal weights; multiplicity = occurrences in the whole solution (as upstream)
        runs = [(w, sol.count(w)) for i, w in enumerate(sol) if i == 0 or sol[i - 1] != w]
        pools = [list(combinations_with_replacement(MASS_NAMES[w], k)) for w, k in runs]
        named.update(tuple(chain.from_iterable(pick)) for pick in product(*pools))
    return MassExplanations(named)


def explain_mass_with_recursion(mass: float, dp_table: DynamicProgrammingTable, max_modifications=np.inf,
                                threshold=None) -> MassExplanations:
    """Table-free enumeration (reference :206-284).  A separate host algorithm the reference keeps for
    cross-checking; it never touches the table and is not a GPU target."""
    rows = dp_table.masses
    weights = [r.mass for r in rows]
    target, thr = _integerise(mass, threshold, dp_table)
    max_len = dp_table.seq.max_len
    cache = {}

    def walk(remaining, start, used_all, used_ind):
        if used_all > max_modifications or used_ind > round(max_len * rows[start].modification_rate):
            return []
        key = (remaining, start)
        if key in cache:
            return cache[key]
        if abs(remaining) <= thr or remaining == 0:
            return [[]]
        if remaining < 0:
            return []
        found = []
        for i in range(start, len(weights)):
            w = weights[i]
            mod = IS_MOD[w]
            tails = walk(remaining - w, i, used_all + 1 if mod else used_all,
                         0 if i != start else (used_ind + 1 if mod else used_ind))
            found.extend([w] + t for t in tails)
        cache[key] = found
        return found

    return convert_nucleotide_masses_to_names(walk(target, 1, 0, 0))
