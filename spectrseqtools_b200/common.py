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
    def __init__(self, *nucleosides):
        self.nucleosides = tuple(sorted(nucleosides))

    def __iter__(self):
        yield from self.nucleosides

    def __len__(self):
        return len(self.nucleosides)

    def __repr__(self):
        return f"{{{','.join(self.nucleosides)}}}"

    def __eq__(self, other):
        return self.nucleosides == other

    def __hash__(self):
        return hash(self.nucleosides)


def calculate_error_threshold(mass1: float, mass2: float, threshold: float) -> float:
    if ERROR_METHOD == "l1_norm":
        return threshold * (mass1 + mass2)
    if ERROR_METHOD == "l2_norm":
        return threshold * ((mass1**2 + mass2**2) ** 0.5)
    raise NotImplementedError("This error method is not implemented.")


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
