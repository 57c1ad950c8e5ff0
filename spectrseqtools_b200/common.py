"""Production wrapper around the enumeration (reference common.py:12-65, without the RAW-file helpers).

``calculate_explanations`` is the one call site through which prediction.py and skeleton_building.py
reach ``explain_mass_with_table``; ``calculate_explanations_batch`` is its batched form.
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence

from .mass_explanation import explain_mass_with_table, explain_masses
from .mass_table import DynamicProgrammingTable

ERROR_METHOD = "l1_norm"
_NUCLEOSIDE_RE = re.compile(r"\d*[ACGU]")


def parse_nucleosides(sequence: str):
    return _NUCLEOSIDE_RE.findall(sequence)


class Explanation:
    """A composition as a name-sorted tuple: iterable, sized, printable as {a,b,c}, equal to another Explanation
    or to a plain tuple with the same names (what the reference's consumers rely on)."""

    __slots__ = ("nucleosides",)

    def __init__(self, *names: str):
        self.nucleosides = tuple(sorted(names))

    def __iter__(self):
        return iter(self.nucleosides)

    def __len__(self) -> int:
        return len(self.nucleosides)

    def __repr__(self) -> str:
        return "{" + ",".join(self.nucleosides) + "}"

    def __eq__(self, other) -> bool:
        return self.nucleosides == (other.nucleosides if isinstance(other, Explanation) else other)

    def __hash__(self) -> int:
        return hash(self.nucleosides)


_ERROR_NORMS = {
    "l1_norm": lambda m1, m2: m1 + m2,
    "l2_norm": lambda m1, m2: (m1**2 + m2**2) ** 0.5,
}


def calculate_error_threshold(mass1: float, mass2: float, threshold: float) -> float:
    """Absolute threshold of a mass difference: relative tolerance times the norm of the two observed masses."""
    norm = _ERROR_NORMS.get(ERROR_METHOD)
    if norm is None:
        raise NotImplementedError("This error method is not implemented.")
    return threshold * norm(mass1, mass2)


def _budget(dp_table) -> int:
    return round(dp_table.seq.modification_rate * dp_table.seq.max_len)


def calculate_explanations(diff: float, threshold: float, dp_table: DynamicProgrammingTable) -> Optional[List[Explanation]]:
    found = explain_mass_with_table(diff, dp_table=dp_table, max_modifications=_budget(dp_table), threshold=threshold).explanations
    if found is None:
        return None
    return [Explanation(*names) for names in found]


def calculate_explanations_batch(diffs: Sequence[float], thresholds: Sequence[float],
                                 dp_table: DynamicProgrammingTable) -> List[Optional[List[Explanation]]]:
    """All (diff, threshold) pairs of a ladder in one device pass; entry p equals calculate_explanations(diffs[p], ...)."""
    batch = explain_masses(diffs, dp_table, max_modifications=_budget(dp_table), thresholds=list(thresholds))
    out: List[Optional[List[Explanation]]] = []
    for p in range(len(batch)):
        found = batch.explanations(p).explanations
        out.append(None if found is None else [Explanation(*names) for names in found])
    return out
