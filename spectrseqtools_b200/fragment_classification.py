"""Fragment classification on the B200 behind the reference's ``spectrseqtools.fragment_classification`` names
(SURVEY §8f row N2).

The reference copies every fragment once per breakage offset, calls ``is_valid_mass`` and ``is_singleton`` from
Python for each copy (``map_elements``, fragment_classification.py:39-82) and filters.  Here the whole
(fragment x breakage) grid is one kernel launch (``sst_classify``): the host ships F observed masses and B
offsets instead of F*B (mass, threshold) pairs, and gets one flag byte per pair back.

``classify_observed`` is the batched array form; ``classify_fragments`` keeps the reference's signature and
column semantics on top of it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _cabi
from .mass_table import DynamicProgrammingTable

MAX_VARIANCE = 1


class ClassifiedMasses:
    """Result of ``classify_observed``: one flag byte per (breakage, fragment), breakage-major."""

    def __init__(self, observed: np.ndarray, breakage_weights: List[int], labels: List[str], precision: float,
                 flags: np.ndarray, pending=None, packed: bool = False):
        self.observed = observed
        self.breakage_weights = breakage_weights
        self.labels = labels          # first label of every weight, like the reference's pl.lit(breakages[0])
        self.precision = precision
        self._flags = flags           # uint8[B, F], or two flags per byte uint8[B, ceil(F / 2)] (unpacked on first use)
        self._packed = packed
        self._pending = pending       # context of an asynchronous call that has not been waited for yet

    def wait(self) -> "ClassifiedMasses":
        if self._pending is not None:
            self._pending.classify_wait()
            self._pending = None
        return self

    @property
    def flags(self) -> np.ndarray:
        self.wait()
        if self._packed:
            F = len(self.observed)
            wide = np.empty((self._flags.shape[0], 2 * self._flags.shape[1]), dtype=np.uint8)
            wide[:, 0::2] = self._flags & 15
            wide[:, 1::2] = self._flags >> 4
            self._flags, self._packed = wide[:, :F], False
        return self._flags

    @property
    def valid(self) -> np.ndarray:
        return (self.flags & _cabi.CLASS_VALID) != 0

    @property
    def out_of_table(self) -> np.ndarray:
        return (self.flags & _cabi.CLASS_OUT_OF_TABLE) != 0

    @property
    def singleton(self) -> np.ndarray:
        return (self.flags & _cabi.CLASS_SINGLETON) != 0

    @property
    def standard_unit_mass(self) -> np.ndarray:
        """observed - breakage_weight * precision, float64[B, F] (the same two IEEE operations as upstream)."""
        off = np.array([w * self.precision for w in self.breakage_weights], dtype=np.float64)
        return self.observed[None, :] - off[:, None]


_OFFSET_CACHE: dict = {}  # the offsets / labels of the last breakage dictionary (same object, same size: same offsets)


def classify_observed(observed: Sequence[float], dp_table: DynamicProgrammingTable, breakage_dict: Dict[int, List[str]],
                      copy: bool = True, wait: bool = True, slot: int = 0, out=None) -> ClassifiedMasses:
    """Validity + singleton flags of every observed mass under every breakage offset: one device pass.

    ``wait=False`` queues the whole call (copies and kernel) on the context's side stream and returns at once; the
    flags are waited for on first access (``.flags`` / ``.wait()``), so an ``explain_masses`` call issued in between
    overlaps it; the flags then cross the bus two per byte.  ``copy=False`` hands out the context's pinned buffer
    (valid until the next classification on the same ``slot``, see ``_cabi.context``).  ``out`` (with ``wait=False``): a
    page-locked uint8 buffer of the caller's for the packed flags (see ``explain_masses(out_block=...)``)."""
    observed = np.ascontiguousarray(observed, dtype=np.float64).reshape(-1)
    # (NaN / infinite masses: the library raises what upstream's int(round(x)) inside is_valid_mass raises)
    key = (id(breakage_dict), len(breakage_dict), dp_table.precision)
    hit = _OFFSET_CACHE.get("last")
    if hit is not None and hit[0] == key and hit[1] is breakage_dict:
        _k, _d, weights, offsets, labels = hit
    else:
        weights = list(breakage_dict.keys())
        offsets = np.array([w * dp_table.precision for w in weights], dtype=np.float64)  # int * float, as upstream
        labels = [breakage_dict[w][0] for w in weights]
        _OFFSET_CACHE["last"] = (key, breakage_dict, weights, offsets, labels)
    dev = dp_table.device_table()
    ctx = dev.ctx if slot == 0 else _cabi.context(dev.ctx.device, slot)
    if not wait:
        flags = ctx.classify_async(dev, observed, offsets, dp_table.precision, dp_table.tolerance, packed=True, out=out)
        return ClassifiedMasses(observed, weights, labels, dp_table.precision, flags, pending=ctx, packed=True)
    ctx.classify_stage(observed, offsets)
    ctx.classify_run(dev, dp_table.precision, dp_table.tolerance)
    flags = ctx.classify_fetch(copy=copy)
    return ClassifiedMasses(observed, weights, labels, dp_table.precision, flags)


def is_singleton(mass, integer_masses, dp_table, threshold=None) -> bool:
    """Reference fragment_classification.py:104-119 (host scalar; the batched form is a bit of ``classify_observed``)."""
    target = int(round(mass / dp_table.precision, 0))
    if threshold is None:
        threshold = dp_table.tolerance * mass
    threshold = int(np.ceil(threshold / dp_table.precision))
    masses = set(integer_masses)
    return any(v in masses for v in range(target - threshold, target + threshold + 1))


def _column(frame, name: str) -> list:
    return list(frame.get_column(name).to_list())


def classify_fragments(fragment_masses, dp_table: DynamicProgrammingTable, breakage_dict: dict, output_file_path=None,
                       intensity_cutoff=0.5e6, mass_cutoff=50000):
    """Reference signature and semantics (fragment_classification.py:17-101): every (fragment, breakage) copy whose
    standard-unit mass is explainable, with ``standard_unit_mass``, ``breakage`` and ``is_singleton`` columns, sorted
    by standard-unit mass, filtered by intensity, observed mass and the sequence mass.  Returns a frame of the
    same kind the alphabet lives in (polars when installed, the stand-in otherwise)."""
    from .masses import _pl as pl

    cols = {c: _column(fragment_masses, c) for c in fragment_masses.columns}
    n = len(next(iter(cols.values()))) if cols else 0
    if "intensity" not in cols:
        cols["intensity"] = [intensity_cutoff * 1.1] * n
    if "neutral_mass" in cols:
        cols = {("observed_mass" if k == "neutral_mass" else k): v for k, v in cols.items()}
    observed = np.array(cols["observed_mass"], dtype=np.float64)
    res = classify_observed(observed, dp_table, breakage_dict)
    if res.out_of_table.any():
        raise _cabi.TableTooSmall("A value of the mass window is not in the DP table. Extend its size if you want to compute larger masses.")
    su = res.standard_unit_mass
    b_idx, f_idx = np.nonzero(res.valid)  # breakage-major, fragments ascending: the reference's concat + filter order
    su_sel = su[b_idx, f_idx]
    intensity = np.array(cols["intensity"], dtype=np.float64)[f_idx]
    keep = (intensity > intensity_cutoff) & (observed[f_idx] < mass_cutoff)
    seq_mass = dp_table.seq.su_mass
    labels = np.array(res.labels, dtype=object)[b_idx]
    complete = np.array(["START" in s and "END" in s for s in labels], dtype=bool)
    keep &= su_sel < seq_mass + MAX_VARIANCE
    keep &= (su_sel > seq_mass - MAX_VARIANCE) | ~complete
    order = np.argsort(su_sel, kind="stable")
    order = order[keep[order]]
    out = {"fragment_index": [int(f_idx[i]) for i in order]}
    for c, v in cols.items():
        out[c] = [v[int(f_idx[i])] for i in order]
    out["standard_unit_mass"] = [float(su_sel[i]) for i in order]
    out["breakage"] = [str(labels[i]) for i in order]
    out["is_singleton"] = [bool(res.singleton[b_idx[i], f_idx[i]]) for i in order]
    frame = pl.DataFrame(out)
    if output_file_path is not None:
        frame.write_csv(output_file_path, separator="\t")
    return frame


def filter_by_sequence_mass(mass_cutoff: float, fragments):
    """Reference fragment_classification.py:122-139 on a frame with ``standard_unit_mass`` and ``breakage``."""
    from .masses import _pl as pl

    su = np.array(_column(fragments, "standard_unit_mass"), dtype=np.float64)
    labels = _column(fragments, "breakage")
    complete = np.array(["START" in s and "END" in s for s in labels], dtype=bool)
    keep = (su < mass_cutoff + MAX_VARIANCE) & ((su > mass_cutoff - MAX_VARIANCE) | ~complete)
    return pl.DataFrame({c: [v for v, k in zip(_column(fragments, c), keep) if k] for c in fragments.columns})
