"""Ladder differences and the explanation-based alphabet reduction, batched.

The reference walks the sorted standard-unit masses of each side with a two-pointer window, calls
``calculate_explanations`` once per pair and once per singleton (prediction.py:261-329), collects the nucleosides
that occur in any explanation, shrinks the alphabet to them — which rebuilds the DP table — re-validates every
fragment against the new table (prediction.py:204-227) and repeats until the alphabet stops shrinking
(prediction.py:170-202).  Here one round is three device passes: ONE enumeration batch for all pairs and
singletons (``sst_explain``), ONE table build (``sst_table_build``) and ONE validity batch (``sst_is_valid``);
the control flow stays on the host, as upstream.

Inputs are plain arrays (the columns ``standard_unit_mass``, ``observed_mass``, ``breakage``, ``is_singleton`` of
the classified-fragments frame); results use the reference's structures (dict diff -> list[Explanation]).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

from . import _cabi
from .common import Explanation, calculate_error_threshold, calculate_explanations_batch
from .mass_explanation import are_valid_masses
from .masses import PHOSPHATE_LINK_MASS
from .mass_table import DynamicProgrammingTable


def ladder_pairs(su_masses: Sequence[float], max_weight: float) -> List[Tuple[int, int]]:
    """(start, end) index pairs in the exact order the reference's window visits them (prediction.py:296-327),
    including its tail behaviour: once ``end`` has reached the last fragment only ``start`` advances."""
    n = len(su_masses)
    pairs: List[Tuple[int, int]] = []
    start, end = 0, 1
    while end < n:
        if end - start <= 0:
            end += 1
            continue
        if su_masses[end] - su_masses[start] > max_weight:
            start += 1
            end = start + 1
            continue
        pairs.append((start, end))
        if end == n - 1:
            start += 1
        else:
            end += 1
    return pairs


def _max_weight(explanation_masses) -> float:
    return max(explanation_masses.get_column("monoisotopic_mass").to_list()) + PHOSPHATE_LINK_MASS


def collect_diff_explanations(su_masses: Sequence[float], observed_masses: Sequence[float], breakages: Sequence[str],
                              is_singleton: Sequence[bool], dp_table: DynamicProgrammingTable,
                              explanation_masses) -> Dict[float, Optional[List[Explanation]]]:
    """``Predictor.collect_diff_explanations_for_su`` (prediction.py:261-284) with one device batch.

    Dict semantics are kept: START side first, then END side, then singletons; a later entry with an equal key
    replaces an earlier one; a pair only enters when it has at least one explanation, a singleton always does."""
    su = np.asarray(su_masses, dtype=np.float64)
    obs = np.asarray(observed_masses, dtype=np.float64)
    max_weight = _max_weight(explanation_masses)
    keys: List[float] = []
    thresholds: List[float] = []
    always: List[bool] = []
    for tag in ("START", "END"):
        idx = [i for i, b in enumerate(breakages) if tag in b]
        side_su = [float(su[i]) for i in idx]
        side_obs = [float(obs[i]) for i in idx]
        for s, e in ladder_pairs(side_su, max_weight):
            keys.append(side_su[e] - side_su[s])
            thresholds.append(calculate_error_threshold(side_obs[s], side_obs[e], dp_table.tolerance))
            always.append(False)
    for i, flag in enumerate(is_singleton):
        if flag:
            keys.append(float(su[i]))
            thresholds.append(dp_table.tolerance * float(obs[i]))
            always.append(True)
    found = calculate_explanations_batch(keys, thresholds, dp_table) if keys else []
    out: Dict[float, Optional[List[Explanation]]] = {}
    for key, expl, keep in zip(keys, found, always):
        if keep or (expl is not None and len(expl) >= 1):
            out[key] = expl
    return out


def observed_nucleotides(explanations: Dict[float, Optional[List[Explanation]]]) -> Set[str]:
    return {nuc for expls in explanations.values() if expls is not None for expl in expls for nuc in expl}


def reduce_alphabet(nucleotides: Set[str], su_masses: Sequence[float], observed_masses: Sequence[float],
                    dp_table: DynamicProgrammingTable) -> np.ndarray:
    """``Predictor._reduce_alphabet`` (prediction.py:204-227): zero the rates of unobserved modifications (device
    table rebuild) and return the keep-mask of the fragments that are still explainable."""
    dp_table.adapt_individual_modification_rates_by_alphabet_reduction(nucleotides)
    su = np.asarray(su_masses, dtype=np.float64)
    obs = np.asarray(observed_masses, dtype=np.float64)
    codes = are_valid_masses(su, dp_table, dp_table.tolerance * obs)
    if (codes == _cabi.VALID_OUT_OF_TABLE).any():
        raise _cabi.TableTooSmall("A value of the mass window is not in the DP table. Extend its size if you want to compute larger masses.")
    return codes == _cabi.VALID_YES


def filter_by_explanation(su_masses: Sequence[float], observed_masses: Sequence[float], breakages: Sequence[str],
                          is_singleton: Sequence[bool], dp_table: DynamicProgrammingTable, explanation_masses):
    """``Predictor.filter_by_explanation`` (prediction.py:170-202): repeat explanation -> reduction until the
    alphabet is stable.  Returns (indices of the surviving fragments, explanations of the last round)."""
    su = np.asarray(su_masses, dtype=np.float64)
    obs = np.asarray(observed_masses, dtype=np.float64)
    brk = list(breakages)
    single = np.asarray(is_singleton, dtype=bool)
    alive = np.arange(len(su))
    old_size = -1
    explanations: Dict[float, Optional[List[Explanation]]] = {}
    while old_size != len(dp_table.masses):
        old_size = len(dp_table.masses)
        explanations = collect_diff_explanations(su[alive], obs[alive], [brk[i] for i in alive], single[alive], dp_table,
                                                 explanation_masses)
        keep = reduce_alphabet(observed_nucleotides(explanations), su[alive], obs[alive], dp_table)
        alive = alive[keep]
    return alive, explanations
