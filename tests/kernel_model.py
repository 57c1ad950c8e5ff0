"""Pure-Python model of the device algorithms (NOT the oracle, NOT the product).

The CUDA enumerator does not walk the table the way the reference does: it uses mass-major row masks
(one mask per mass instead of ~R dependent UP reads) and, for the first-visit memo semantics, a
mass-at-a-time replay (phase A) followed by plain path enumeration (phase B).  This file states those
algorithms in Python with the same structure as csrc/sst_explain.cuh so that their LOGIC can be checked
against the golden vectors on a machine without a GPU (tests/test_kernel_model.py).  The GPU tests then
check the kernels themselves.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

INF = 1 << 30


def row_masks(table: np.ndarray) -> List[int]:
    """H[v]: Python int with bit r set iff bit1(r, v) in the packed table (rows 1..R-1)."""
    R, C = table.shape
    n = C * 32
    H = [0] * n
    for r in range(1, R):
        row = table[r]
        for j in range(C):
            word = int(row[j])
            if not word & 0xAAAAAAAAAAAAAAAA:
                continue
            for k in range(32):
                if (word >> (2 * (31 - k) + 1)) & 1:
                    H[j * 32 + k] |= 1 << r
    return H


def last_row_reach(table: np.ndarray, v: int) -> bool:
    word = int(table[-1, v // 32])
    return bool((word >> (2 * (31 - v % 32))) & 3)


def _low(r: int) -> int:  # rows <= r
    return (1 << (r + 1)) - 1


def phase_a(H: Sequence[int], table: np.ndarray, weights: Sequence[int], is_mod: Sequence[bool], ind: Sequence[int],
            lo: int, hi: int, max_mods: int) -> Dict[int, int]:
    """First-visit replay: returns {mass: alive-edge mask}.  Mirrors k_memo_phase_a."""
    R = len(weights)
    limit = table.shape[1] * 32
    top: Dict[int, int] = {}
    alive: Dict[int, int] = {}

    def arrive(m: int, r_in: int, all_left: int, ind_left: int) -> bool:
        t = top.get(m, 0)
        if m not in alive:
            alive[m] = 0
        if t >= r_in:
            return bool(alive[m] & _low(r_in))
        pend = H[m] & _low(r_in) & ~_low(t)
        new = 0
        r = 1
        while pend:
            if pend & 1 << r:
                pend &= ~(1 << r)
                ind_here = ind_left if r == r_in else ind[r]
                mod = 1 if is_mod[r] else 0
                if not (mod and not (all_left > 0 and ind_here > 0)):
                    m2 = m - weights[r]
                    if m2 == 0 or arrive(m2, r, all_left - mod, ind_here - mod):
                        new |= 1 << r
            r += 1
        alive[m] |= new
        top[m] = r_in
        return bool(alive[m] & _low(r_in))

    for v in range(max(lo, 1), min(hi, limit - 1) + 1):
        if last_row_reach(table, v):
            arrive(v, R - 1, max_mods, ind[R - 1])
    return alive


def enumerate_root(v: int, mask_of, weights: Sequence[int], is_mod: Sequence[bool], ind: Sequence[int], max_mods: int,
                   exact: bool) -> List[List[int]]:
    """All compositions (ascending row lists) under root value v.  Mirrors k_enumerate."""
    R = len(weights)
    out: List[List[int]] = []
    path: List[int] = []

    def rec(m: int, rmax: int, all_left: int, ind_left: int):
        mask = mask_of(m) & _low(rmax)
        r = 1
        while mask:
            if mask & 1 << r:
                mask &= ~(1 << r)
                ca, ci = 0, 0
                ok = True
                if exact:
                    ind_here = ind_left if r == rmax else ind[r]
                    mod = 1 if is_mod[r] else 0
                    if mod and not (all_left > 0 and ind_here > 0):
                        ok = False
                    ca, ci = all_left - mod, ind_here - mod
                if ok:
                    m2 = m - weights[r]
                    path.append(r)
                    if m2 == 0:
                        out.append(list(reversed(path)))
                    else:
                        rec(m2, r, ca, ci)
                    path.pop()
            r += 1

    rec(v, R - 1, max_mods, ind[R - 1])
    return out


def budgets_cannot_bind(weights, is_mod, ind, max_mods, hi: int) -> bool:
    """Host-side test for the FREE mode (same rule as spectrseqtools_b200.mass_explanation)."""
    mods = [r for r in range(1, len(weights)) if is_mod[r]]
    if not mods or hi <= 0:
        return True
    if max_mods < hi // min(weights[r] for r in mods):
        return False
    return all(ind[r] >= hi // weights[r] for r in mods)


def explain(table: np.ndarray, weights, is_mod, ind, target: int, thr: int, max_mods: int, with_memo: bool):
    """-> (solutions as ascending row lists in device order, has_empty_solution)."""
    R = len(weights)
    limit = table.shape[1] * 32
    lo, hi = target - thr, target + thr
    if lo <= hi and hi >= limit:
        raise NotImplementedError("out of table")
    H = row_masks(table)
    if budgets_cannot_bind(weights, is_mod, ind, max_mods, hi):
        mode = "free"
    else:
        mode = "memo" if with_memo else "exact"
    if mode == "memo":
        alive = phase_a(H, table, weights, is_mod, ind, lo, hi, max_mods)
        mask_of = lambda m: alive.get(m, 0)  # noqa: E731
    else:
        mask_of = lambda m: H[m]  # noqa: E731
    sols: List[List[int]] = []
    for v in range(max(lo, 1), min(hi, limit - 1) + 1):
        if last_row_reach(table, v):
            sols += enumerate_root(v, mask_of, weights, is_mod, ind, max_mods, mode == "exact")
    return sols, (lo <= 0 <= hi), mode


# ---------------------------------------------------------------- N3: the ladder window as count -> scan -> fill
def ladder_pairs_closed_form(su: Sequence[float], max_weight: float):
    """The pairs of ``Predictor.collect_explanations_per_side`` (reference prediction.py:296-327) the way
    csrc/sst_ladder.cuh makes them: E(s) = last e with su[e] - su[s] <= max_weight (binary search, sorted masses), s0 =
    first s whose E(s) is the last fragment; s <= s0 pairs with s+1 .. E(s), every later s with the last fragment only
    (once ``end`` sits on the last fragment the reference only moves ``start``)."""
    n = len(su)
    reach = []
    for s in range(n):
        lo, hi = s, n - 1
        while lo < hi:
            mid = (lo + hi + 1) >> 1
            if not (su[mid] - su[s] > max_weight):
                lo = mid
            else:
                hi = mid - 1
        reach.append(lo)
    s0 = min([s for s in range(n) if reach[s] == n - 1 and reach[s] > s], default=n + 1)
    pairs = []
    for s in range(n):
        if s <= s0:
            pairs += [(s, e) for e in range(s + 1, reach[s] + 1)]
        elif s < n - 1:
            pairs.append((s, n - 1))
    return pairs
