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


class DeviceLadder:
    """The classified-fragment frame on the device (rows N3 + N4): pairs, thresholds, enumeration, dedup by key, the
    union of observed table rows and the re-validation all run there; per round the host sees a 128-bit row mask and
    two counters (``sst_ladder_*``, kernels in csrc/sst_ladder.cuh).

    ``su_masses`` must ascend (prediction.py:68-72 sorts the frame before ``filter_by_explanation``)."""

    def __init__(self, su_masses: Sequence[float], observed_masses: Sequence[float], breakages: Sequence[str],
                 is_singleton: Sequence[bool], dp_table: DynamicProgrammingTable, explanation_masses):
        self.dp = dp_table
        self.max_weight = _max_weight(explanation_masses)
        su = np.ascontiguousarray(su_masses, dtype=np.float64)
        flags = np.array([(1 if "START" in b else 0) | (2 if "END" in b else 0) | (4 if s else 0)
                          for b, s in zip(breakages, is_singleton)], dtype=np.uint8)
        self.ctx = dp_table.device_table().ctx
        self.ctx.ladder_stage(su, np.ascontiguousarray(observed_masses, dtype=np.float64), flags)
        self.n_calls = self.n_compositions = 0

    def round(self) -> Set[str]:
        """One ``collect_diff_explanations_for_su`` over the alive fragments -> the nucleosides that occur in any
        surviving explanation (every name of every table row used, as ``convert_nucleotide_masses_to_names`` expands)."""
        from .common import _budget
        from .mass_explanation import MASS_NAMES, _row_metadata

        dp = self.dp
        dev = dp.device_table()
        weights, is_mod, ind = _row_metadata(dp)
        mask, self.n_calls, self.n_compositions = self.ctx.ladder_round(dev, self.max_weight, dp.precision, dp.tolerance, _budget(dp), ind,
                                                                        is_mod, True)
        self._weights = weights
        return {name for r in range(1, len(weights)) if (mask >> r) & 1 for name in MASS_NAMES[int(weights[r])]}

    def revalidate(self) -> int:
        """``Predictor._reduce_alphabet``'s loop over the fragments against the current (rebuilt) table -> fragments alive."""
        return self.ctx.ladder_revalidate(self.dp.device_table(), self.dp.precision, self.dp.tolerance)

    def alive(self) -> np.ndarray:
        return self.ctx.ladder_fetch(calls=False)[0].astype(bool)

    def calls(self):
        """(keys, thresholds, flags) of the last round in generation order (START pairs, END pairs, singletons);
        flag bit 0: the call entered its dict, bit 1: it is the entry that survives under its key."""
        _alive, keys, thr, fl = self.ctx.ladder_fetch()
        return keys, thr, fl

    def explanations(self) -> Dict[float, Optional[List[Explanation]]]:
        """The dict ``collect_diff_explanations_for_su`` returns for the last round (surviving entries only are read back)."""
        from .mass_explanation import ExplanationBatch

        keys, _thr, fl = self.calls()
        out: Dict[float, Optional[List[Explanation]]] = {}
        if not len(keys):
            return out
        status, off, recs = self.ctx.explain_fetch(copy=False)
        batch = ExplanationBatch(status, off, recs, self._weights, [m.names for m in self.dp.masses])
        order: Dict[float, int] = {}
        for i in np.nonzero(fl & 1)[0]:  # key order = first insertion, value = the last entering call's
            order.setdefault(float(keys[i]), 0)
        for i in np.nonzero(fl & 2)[0]:
            found = batch.explanations(int(i)).explanations
            order[float(keys[i])] = None if found is None else [Explanation(*names) for names in found]
        out.update(order)
        return out


def filter_by_explanation_device(su_masses: Sequence[float], observed_masses: Sequence[float], breakages: Sequence[str],
                                 is_singleton: Sequence[bool], dp_table: DynamicProgrammingTable, explanation_masses):
    """``Predictor.filter_by_explanation`` (prediction.py:170-202) with the fragment frame resident on the device: per
    round one generated-and-enumerated batch, one table rebuild, one re-validation; the host only turns the 128-bit row
    mask into the reduced alphabet.  Returns (indices of the surviving fragments, explanations of the last round)."""
    lad = DeviceLadder(su_masses, observed_masses, breakages, is_singleton, dp_table, explanation_masses)
    old_size = -1
    while old_size != len(dp_table.masses):
        old_size = len(dp_table.masses)
        nucleotides = lad.round()
        dp_table.adapt_individual_modification_rates_by_alphabet_reduction(nucleotides)
        lad.revalidate()
    # the last round left the alphabet — and with it the table and its row numbering — as it was: its explanations
    # are still on the device
    return np.nonzero(lad.alive())[0], lad.explanations()


def filter_by_explanation(su_masses: Sequence[float], observed_masses: Sequence[float], breakages: Sequence[str],
                          is_singleton: Sequence[bool], dp_table: DynamicProgrammingTable, explanation_masses):
    """``Predictor.filter_by_explanation`` (prediction.py:170-202): repeat explanation -> reduction until the
    alphabet is stable.  Returns (indices of the surviving fragments, explanations of the last round).  (Host loop
    over batched device calls; ``filter_by_explanation_device`` keeps the frame on the device.)"""
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
