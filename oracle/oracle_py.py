"""TEST INFRASTRUCTURE — CPU restatement (numpy / pure Python) of the reference's mass-explanation path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module, and only as the checker or the timed CPU baseline.  The product package
``spectrseqtools_b200`` never imports it.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the reference's unmodified function bodies (via
``oracle/ref_harness.py``) in the build container and commits their outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement (and the C one in ``oracle/oracle.c``) against them.

Every function cites the reference lines it follows (paths relative to /root/reference/spectrseqtools).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from itertools import chain, combinations_with_replacement, product
from typing import Dict, List, Optional, Sequence

import numpy as np

_CELL_TYPES = {4: np.uint8, 8: np.uint16, 16: np.uint32, 32: np.uint64}


class OutOfTable(NotImplementedError):
    """A probed integer mass lies beyond the table (mass_explanation.py:69-73,134-138)."""


@dataclass
class Row:
    """One table row: integer weight + budget metadata (mass_table.py:29-34)."""

    mass: int
    is_modification: bool
    modification_rate: float


def integerise(mass: float, threshold: Optional[float], precision: float, tolerance: float):
    """(target, integer threshold) exactly as mass_explanation.py:51-58 / :107-114 compute them."""
    target = int(round(mass / precision, 0))
    if threshold is None:
        threshold = tolerance * mass
    thr = int(np.ceil(threshold / precision))
    return target, thr


def last_column_mask(max_mass: int, compression: int) -> int:
    """The ``full << 2*(max_col - (max_mass+1) % max_col)`` of mass_table.py:246, numpy semantics kept."""
    ctype = _CELL_TYPES[compression]
    max_col = int(np.ceil((max_mass + 1) / compression))
    full = ctype(np.iinfo(ctype).max)
    shift = 2 * (max_col - (max_mass + 1) % max_col)
    if shift > np.iinfo(ctype).max:
        # numpy >= 2 refuses to turn the Python int into the cell dtype; the reference dies the same way
        raise OverflowError(f"Python integer {shift} out of bounds for {np.dtype(ctype).name}")
    return int(full << shift) if shift < 8 * np.dtype(ctype).itemsize else 0  # numpy: shift >= width -> 0


def build_bit_table_closed_form(weights: Sequence[int], max_mass: int, compression: int = 32) -> np.ndarray:
    """Vectorised numpy statement of what mass_table.py:207-248 produces (valid when every weight >= compression).

    bit0(i, v) = reach_{i-1}[v], bit1(i, v) = reach_i[v - w_i] with reach_0 = {0} and
    reach_i = reach_{i-1} U (reach_i + w_i); cells beyond max_mass cleared by the last-column mask.
    Works on an unpacked boolean reach vector, then packs MSB-first.  The literal word-by-word loop
    lives in oracle.c (``oracle_build_bit_table``); the two are compared in tests/test_oracle.py.
    """
    ctype = _CELL_TYPES[compression]
    R = len(weights)
    max_col = int(np.ceil((max_mass + 1) / compression))
    n = max_col * compression
    reach = np.zeros(n, dtype=bool)
    reach[0] = True
    cells = np.zeros((R, n), dtype=np.uint8)
    cells[0, 0] = 3
    for i in range(1, R):
        w = int(weights[i])
        if w < compression:
            raise ValueError("closed form needs weights >= compression")
        prev = reach.copy()
        # unbounded knapsack on one weight: sweep slabs of width w in ascending order
        for start in range(w, n, w):
            stop = min(start + w, n)
            reach[start:stop] |= reach[start - w : stop - w]
        bit1 = np.zeros(n, dtype=bool)
        bit1[w:] = reach[: n - w]
        cells[i] = prev.astype(np.uint8) | (bit1.astype(np.uint8) << 1)
    packed = np.zeros((R, max_col), dtype=ctype)
    view = cells.reshape(R, max_col, compression)
    for k in range(compression):
        packed |= view[:, :, k].astype(ctype) << ctype(2 * (compression - 1 - k))
    packed[:, -1] &= ctype(last_column_mask(max_mass, compression))
    return packed


def build_mass_table(weights: Sequence[int], max_mass: int) -> np.ndarray:
    """Byte-per-mass table, mass_table.py:292-316: cell = 1 if reachable with earlier rows, + 2 if (mass - w_i) is a
    non-zero cell of the same row (ascending in-place sweep = unbounded use of the row); cell (0, 0) = 3."""
    R = len(weights)
    t = np.zeros((R, max_mass + 1), dtype=np.uint8)
    t[0, 0] = 3
    for i in range(1, R):
        w = int(weights[i])
        t[i] = (t[i - 1] != 0).astype(np.uint8)
        for start in range(0, max_mass + 1 - w, w):  # slabs of width w: sources of a slab are final before it is read
            stop = min(start + w, max_mass + 1 - w)
            t[i, start + w : stop + w] += 2 * (t[i, start:stop] != 0).astype(np.uint8)
    return t


def _cell(table: np.ndarray, row: int, mass: int, compression: int) -> int:
    """2-bit-aligned shifted cell value, mass_explanation.py:140-145 (low bits = cell of `mass`)."""
    return int(table[row, mass // compression]) >> (2 * (compression - 1 - mass % compression))


def is_valid_mass(mass: float, table: np.ndarray, compression: int, precision: float, tolerance: float,
                  threshold: Optional[float] = None) -> bool:
    """mass_explanation.py:45-89."""
    target, thr = integerise(mass, threshold, precision, tolerance)
    last = len(table) - 1
    limit = len(table[0]) * compression
    for value in range(target - thr, target + thr + 1):
        if value <= 0:
            continue
        if value >= limit:
            raise OutOfTable(f"The value {value} is not in the DP table.")
        cell = _cell(table, last, value, compression)
        if cell % compression == 0:
            continue
        if cell & 3:
            return True
    return False


def is_singleton(mass: float, integer_masses: Sequence[int], precision: float, tolerance: float,
                 threshold: Optional[float] = None) -> bool:
    """fragment_classification.py:104-119: some row weight (the leading 0 included) lies inside the window."""
    target, thr = integerise(mass, threshold, precision, tolerance)
    for value in range(target - thr, target + thr + 1):
        if value in integer_masses:
            return True
    return False


def classify_pairs(observed: Sequence[float], breakage_weights: Sequence[int], table: np.ndarray, integer_masses: Sequence[int],
                   compression: int, precision: float, tolerance: float) -> np.ndarray:
    """fragment_classification.py:39-82 as flags[b][f]: bit 1 valid, bit 2 out-of-table before any hit, bit 4
    singleton.  standard-unit mass = observed - (weight * precision), threshold = tolerance * observed."""
    out = np.zeros((len(breakage_weights), len(observed)), dtype=np.uint8)
    for b, bw in enumerate(breakage_weights):
        for f, x in enumerate(observed):
            su = x - (bw * precision)
            thr = tolerance * x
            try:
                code = 1 if is_valid_mass(su, table, compression, precision, tolerance, thr) else 0
            except OutOfTable:
                code = 2
            if is_singleton(su, integer_masses, precision, tolerance, thr):
                code |= 4
            out[b, f] = code
    return out


def individual_budgets(rows: Sequence[Row], max_len: int) -> List[int]:
    """IND[r] = round(max_len * rate_r), Python banker's rounding (mass_explanation.py:158-161,200)."""
    return [round(max_len * r.modification_rate) for r in rows]


def explain_solutions(mass: float, table: np.ndarray, rows: Sequence[Row], max_len: int, compression: int,
                      precision: float, tolerance: float, max_modifications=math.inf,
                      threshold: Optional[float] = None, with_memo: bool = True) -> List[List[int]]:
    """The list of weight lists built by mass_explanation.py:92-201 (before name conversion).

    Same traversal, same memo keyed (mass, row) WITHOUT budgets (first visit wins), same checks in the
    same order.  Recursion depth is ~R + L, well inside CPython's default limit for the sizes tested.
    """
    target, thr = integerise(mass, threshold, precision, tolerance)
    limit = len(table[0]) * compression
    ind = individual_budgets(rows, max_len)
    memo: Dict[tuple, list] = {}

    def visit(m: int, r: int, all_left, ind_left) -> list:
        if with_memo and (m, r) in memo:
            return memo[(m, r)]
        if m < 0:
            return []
        if m == 0:
            return [[]]
        if m >= limit:
            raise OutOfTable(f"The value {m} is not in the DP table.")
        cell = _cell(table, r, m, compression)
        if cell % compression == 0:
            return []
        found: list = []
        if cell & 1:  # UP: same mass, previous row, fresh individual budget
            found += visit(m, r - 1, all_left, ind[r - 1])
        if cell & 2:  # LEFT: spend one copy of this row's weight
            row = rows[r]
            if not row.is_modification or (all_left > 0 and ind_left > 0):
                if row.is_modification:
                    all_left -= 1
                    ind_left -= 1
                found += [tail + [row.mass] for tail in visit(m - row.mass, r, all_left, ind_left)]
        if with_memo:
            memo[(m, r)] = found
        return found

    out: list = []
    top = len(rows) - 1
    for value in range(target - thr, target + thr + 1):
        out += visit(value, top, max_modifications, ind[top])
    return out


def solutions_to_names(solutions: List[List[int]], mass_names: Dict[int, List[str]]):
    """mass_explanation.py:287-320: None if no solution at all; skip the empty solution; expand names."""
    if len(solutions) == 0:
        return None
    named = set()
    for sol in solutions:
        if not sol:
            continue
        runs = [sol[i] for i in range(len(sol)) if i == 0 or sol[i - 1] != sol[i]]
        pools = [list(combinations_with_replacement(mass_names[m], sol.count(m))) for m in runs]
        named.update(tuple(chain.from_iterable(pick)) for pick in product(*pools))
    return named


def explain_mass_with_recursion(mass: float, rows: Sequence[Row], max_len: int, precision: float, tolerance: float,
                                is_mod: Dict[int, bool], max_modifications=math.inf,
                                threshold: Optional[float] = None) -> List[List[int]]:
    """Table-free variant, mass_explanation.py:206-282 (solutions before name conversion)."""
    weights = [r.mass for r in rows]
    target, thr = integerise(mass, threshold, precision, tolerance)
    memo: Dict[tuple, list] = {}

    def dp(remaining, start, used_all, used_ind):
        if used_all > max_modifications or used_ind > round(max_len * rows[start].modification_rate):
            return []
        if (remaining, start) in memo:
            return memo[(remaining, start)]
        if abs(remaining) <= thr:
            return [[]]
        if remaining == 0:
            return [[]]
        if remaining < 0:
            return []
        combos = []
        for i in range(start, len(weights)):
            w = weights[i]
            mod = is_mod[w]
            sub = dp(remaining - w, i, used_all + 1 if mod else used_all,
                     0 if i != start else (used_ind + 1 if mod else used_ind))
            for c in sub:
                combos.append([w] + c)
        memo[(remaining, start)] = combos
        return combos

    return dp(target, 1, 0, 0)


def sequence_length_bound(table: np.ndarray, rows: Sequence[Row], max_len: int, su_mass: float, obs_mass: float,
                          modification_rate: float, compression: int, precision: float, tolerance: float,
                          direction: str) -> int:
    """mass_table.py:343-487 ("lower" / "upper" nucleotide-count bound over the window)."""
    if direction not in ("lower", "upper"):
        raise NotImplementedError(f"Support for '{direction}' is currently not given.")
    max_mods = round(modification_rate * max_len)
    target = int(round(su_mass / precision, 0))
    thr = int(np.ceil(tolerance * obs_mass / precision))
    default = max_len + 1 if direction == "lower" else -1
    pick = min if direction == "lower" else max
    limit = len(table[0]) * compression
    ind = individual_budgets(rows, max_len)
    memo: Dict[tuple, int] = {}

    def visit(m, r, all_left, ind_left):
        if (m, r) in memo:
            return memo[(m, r)]
        if m < 0:
            return default
        if m == 0:
            return 0
        if m >= limit:
            raise OutOfTable(f"The value {m} is not in the DP table.")
        cell = _cell(table, r, m, compression)
        if cell % compression == 0:
            return default
        bounds = [default]
        if cell & 1:
            bounds.append(visit(m, r - 1, all_left, ind[r - 1]))
        if cell & 2:
            row = rows[r]
            if not row.is_modification or (all_left > 0 and ind_left > 0):
                if row.is_modification:
                    all_left -= 1
                    ind_left -= 1
                bounds.append(visit(m - row.mass, r, all_left, ind_left) + 1)
        memo[(m, r)] = pick(bounds)
        return memo[(m, r)]

    top = len(rows) - 1
    per_value = [visit(v, top, max_mods, ind[top]) for v in range(target - thr, target + thr + 1)]
    best = pick(per_value)
    if best == default:
        best = 1 if direction == "lower" else max_len
    return best


def canonical_digest(named) -> str:
    """First 16 hex of sha256 over the sorted name tuples; 'None' for no solution."""
    import hashlib

    if named is None:
        return "None"
    text = "\n".join(",".join(t) for t in sorted(named))
    return hashlib.sha256(text.encode()).hexdigest()[:16]
