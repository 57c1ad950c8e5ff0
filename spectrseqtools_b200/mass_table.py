"""Device-backed DP table behind the reference's ``spectrseqtools.mass_table`` names.

Same public surface as the reference module (``TABLE_DIR``, ``MAX_SEQ_LENGTH``, ``SequenceInformation``,
``NucleotideMass``, ``DynamicProgrammingTable``, ``set_table_path``, ``initialize_nucleotide_masses``,
``set_up_bit_table``, ``select_table_building_settings``, ``set_up_mass_table``, ``load_dp_table``,
``compute_sequence_length_bound``; reference mass_table.py:14-487).  The table itself is built by the
sm_100a kernel behind ``sst_table_build`` and lives on the GPU; ``DynamicProgrammingTable.table`` copies
it to the host (bit-identical to the reference's ndarray) only when somebody reads it.

There is no CPU build path: without libsst_b200.so and a B200 the constructors raise
``_cabi.DeviceUnavailable``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from .masses import EXPLANATION_MASSES, UNMODIFIED_BASES

try:
    from platformdirs import user_cache_dir

    TABLE_DIR = user_cache_dir(appname="spectrseqtools/dp_table", version="1.3", ensure_exists=False)
except Exception:  # pragma: no cover
    TABLE_DIR = os.path.join(os.path.expanduser("~"), ".cache", "spectrseqtools", "dp_table", "1.3")

# longest sequence (in copies of the heaviest nucleotide) the table covers
MAX_SEQ_LENGTH = 35


@dataclass
class SequenceInformation:
    max_len: int
    su_mass: float
    obs_mass: float
    modification_rate: float


_ROW_VERSION = 0  # bumped by every assignment to a row's mass or modification rate


def row_version() -> int:
    return _ROW_VERSION


@dataclass
class NucleotideMass:
    mass: int
    names: List[str]
    is_modification: bool
    modification_rate: float

    def __setattr__(self, name, value):
        # callers zero and restore rates in place (prediction.py, the alphabet reduction): a change bumps a module-wide
        # version so that per-table caches of derived arrays (mass_explanation._row_metadata) notice without reading
        # every row's rate on every call
        if name == "modification_rate" or name == "mass":
            global _ROW_VERSION
            _ROW_VERSION += 1
        object.__setattr__(self, name, value)

    def __eq__(self, other):
        return self.mass == other.mass

    def __le__(self, other):
        return self.mass <= other.mass

    def __lt__(self, other):
        return self.mass < other.mass

    def __ge__(self, other):
        return self.mass >= other.mass

    def __gt__(self, other):
        return self.mass > other.mass


_CELL_TYPES = {4: np.uint8, 8: np.uint16, 16: np.uint32, 32: np.uint64}


def select_table_building_settings(compression_rate: int):
    """Cell dtype and bit patterns per compression rate (reference mass_table.py:251-289)."""
    if compression_rate not in _CELL_TYPES:
        raise ValueError(f"The compression rate {compression_rate} is not compatible with the table setup.")
    ctype = _CELL_TYPES[compression_rate]
    full = int(np.iinfo(ctype).max)
    return {
        "type": ctype,
        "init": 3 << (2 * (compression_rate - 1)),
        "alt_first": full // 3 * 2,  # 0xAA..: the "one more copy of this row" bit of every cell
        "alt_sec": full // 3,        # 0x55..: the "reachable with earlier rows" bit of every cell
        "full": ctype(full),
    }


def _last_column_mask(max_mass: int, compression_rate: int) -> int:
    """numpy's value of ``full << 2*(max_col - (max_mass+1) % max_col)`` (reference mass_table.py:246).

    Evaluated with numpy scalars on purpose: a shift >= the cell width gives 0 there (the whole last
    word of every row is wiped when (max_mass+1) is a multiple of the compression rate).
    """
    settings = select_table_building_settings(compression_rate)
    max_col = int(np.ceil((max_mass + 1) / compression_rate))
    return int(settings["full"] << 2 * (max_col - (max_mass + 1) % max_col))


# ---- device table cache: one build per (weights, max_mass, compression, device) ------------------
_TABLE_CACHE: Dict[Tuple, "_cabi.DeviceTable"] = {}
_TABLE_CACHE_MAX = 4


def device_table(integer_masses: Sequence[int], max_mass: int, compression_rate: int = 32,
                 device: Optional[int] = None) -> "_cabi.DeviceTable":
    """Build (or reuse) the device-resident table for an alphabet.  GPU only."""
    if compression_rate not in _CELL_TYPES:
        raise ValueError(f"The compression rate {compression_rate} is not compatible with the table setup.")
    ctx = _cabi.context(device)
    key = (ctx.device, tuple(int(m) for m in integer_masses), int(max_mass), int(compression_rate))
    hit = _TABLE_CACHE.get(key)
    if hit is not None:
        return hit
    mask = _last_column_mask(int(max_mass), compression_rate)
    tab = ctx.build_table(key[1], int(max_mass), compression_rate, mask, with_masks=True)
    while len(_TABLE_CACHE) >= _TABLE_CACHE_MAX:
        _TABLE_CACHE.pop(next(iter(_TABLE_CACHE)))
    _TABLE_CACHE[key] = tab
    return tab


def clear_table_cache():
    _TABLE_CACHE.clear()


def set_up_bit_table(integer_masses, max_mass: int, compression_rate: int):
    """2-bit reachability table as a host ndarray, bit-identical to the reference (mass_table.py:207-248).

    Built on the GPU and copied back; callers that only need the table for explanation calls should use
    ``DynamicProgrammingTable`` (no copy).  The device layout is 32 masses per ``uint64`` cell — the only one the
    reference ever uses (``COMPRESSION_RATE``); the narrower layouts it also accepts (4 / 8 / 16 masses per
    ``uint8`` / ``uint16`` / ``uint32`` cell) are re-packed on the host from a slightly wider device table, with that
    layout's own last-column mask.  ``DynamicProgrammingTable`` itself is 32-per-cell only.
    """
    if compression_rate == 32:
        return device_table(integer_masses, max_mass, compression_rate).download()
    settings = select_table_building_settings(compression_rate)  # ValueError for anything but 4 / 8 / 16 / 32
    weights = [int(w) for w in integer_masses]
    if any(w < compression_rate for w in weights[1:]):
        raise ValueError(f"weights below the compression rate ({compression_rate}): the reference's in-place word loop is not a closed form there")
    mask = _last_column_mask(int(max_mass), compression_rate)    # numpy semantics (may raise OverflowError, as upstream)
    ctype = settings["type"]
    max_col = int(np.ceil((max_mass + 1) / compression_rate))
    wide = device_table(weights, int(max_mass) + 64, 32).download()  # no masked cell below max_col * compression_rate
    n = max_col * compression_rate
    cells = np.zeros((wide.shape[0], wide.shape[1] * 32), dtype=np.uint8)
    for k in range(32):
        cells[:, k::32] = ((wide >> np.uint64(2 * (31 - k))) & np.uint64(3)).astype(np.uint8)
    view = cells[:, :n].reshape(wide.shape[0], max_col, compression_rate)
    out = np.zeros((wide.shape[0], max_col), dtype=ctype)
    for k in range(compression_rate):
        out |= view[:, :, k].astype(ctype) << ctype(2 * (compression_rate - 1 - k))
    out[:, -1] &= ctype(mask)
    return out


def set_up_mass_table(integer_masses, max_mass):
    """Byte-per-mass variant (reference mass_table.py:292-316): value 1 = reachable with earlier rows,
    +2 = one more copy of this row.  Derived from the packed device table (no separate kernel), which is built 64
    masses wider than asked: the packed layout masks its last word (:246) — wiping it entirely when max_mass + 1 is
    a multiple of 32 — while the byte table keeps every cell up to max_mass."""
    packed = set_up_bit_table(list(integer_masses), max_mass + 64, 32)
    R, C = packed.shape
    out = np.zeros((R, C * 32), dtype=np.uint8)
    for k in range(32):
        out[:, k::32] = ((packed >> np.uint64(2 * (31 - k))) & np.uint64(3)).astype(np.uint8)
    return np.ascontiguousarray(out[:, : max_mass + 1])


def set_table_path(precision, compression_rate):
    path = f"{TABLE_DIR}/tol_{precision:.0E}.{compression_rate}_per_cell"
    os.makedirs(os.path.dirname(path), exist_ok=True)
    return path


def load_dp_table(table_path, integer_masses):
    """Reference signature (mass_table.py:319-340).  The GPU rebuild is faster than reading 582 MB from
    disk, so nothing is cached on disk; the path only carries the compression rate, as upstream."""
    compression_rate = int(table_path.split(".")[-1].removesuffix("_per_cell"))
    max_mass = max(integer_masses) * MAX_SEQ_LENGTH
    if compression_rate == 1:
        return set_up_mass_table(integer_masses, max_mass)
    return set_up_bit_table(integer_masses, max_mass, compression_rate)


def initialize_nucleotide_masses(nucleotide_df) -> List[NucleotideMass]:
    """Sorted unique integer masses with a leading 0 row (reference mass_table.py:154-204)."""
    ims = list(nucleotide_df.get_column("tolerated_integer_masses").to_list())
    reps = list(nucleotide_df.get_column("nucleoside").to_list())
    rates = list(nucleotide_df.get_column("modification_rate").to_list())
    names: Dict[int, List[str]] = {}
    best_rate: Dict[int, float] = {}
    for m, rep, rate in zip(ims, reps, rates):
        names.setdefault(m, []).append(rep)
        best_rate[m] = max(best_rate.get(m, rate), rate)
    out = []
    for m in sorted(set(ims) | {0}):
        if m == 0:
            out.append(NucleotideMass(0, [], False, 0.0))
        else:
            out.append(NucleotideMass(m, names[m], any(n not in UNMODIFIED_BASES for n in names[m]), best_rate[m]))
    return out


class DynamicProgrammingTable:
    """Holder with the reference's attributes (``table, compression_per_cell, precision, tolerance, seq,
    masses``; mass_table.py:52-139).  ``table`` is materialised lazily from the device."""

    def __init__(self, nucleotide_df, compression_rate: int, tolerance: float, precision: float,
                 seq: SequenceInformation, device: Optional[int] = None):
        self.compression_per_cell = compression_rate
        self.tolerance = tolerance
        self.precision = precision
        self.seq = seq
        self.masses = initialize_nucleotide_masses(nucleotide_df)
        self._device = device
        self._dev: Optional[_cabi.DeviceTable] = None
        self._host: Optional[np.ndarray] = None
        self._adapt_individual_modification_rates_by_universal_one()
        if self._dev is None and self._host is None:
            self._build_device_table()

    # -- the public ndarray, bit-identical to the reference's
    @property
    def table(self) -> np.ndarray:
        if self._host is None:
            self._host = self.device_table().download()
        return self._host

    @table.setter
    def table(self, value):
        # a caller-supplied table replaces ours; it is uploaded on the next explanation call
        self._host = None if value is None else np.asarray(value)
        self._dev = None

    def device_table(self) -> "_cabi.DeviceTable":
        if self._dev is None:
            if self._host is not None:
                self._dev = _cabi.context(self._device).upload_table(self._host, [m.mass for m in self.masses])
            else:
                self._build_device_table()
        return self._dev

    def _build_device_table(self):
        if self.compression_per_cell == 1:
            raise NotImplementedError("the byte-per-mass table (compression 1) has no device path")
        ims = [m.mass for m in self.masses]
        self._dev = device_table(ims, max(ims) * MAX_SEQ_LENGTH, self.compression_per_cell, self._device)
        self._host = None

    # -- alphabet handling (reference mass_table.py:86-121)
    def _adapt_individual_modification_rates_by_universal_one(self):
        for nm in self.masses:
            if nm.is_modification and nm.modification_rate > self.seq.modification_rate:
                nm.modification_rate = self.seq.modification_rate
        self._reduce_nucleotide_list()

    def adapt_individual_modification_rates_by_alphabet_reduction(self, alphabet):
        for nm in self.masses:
            if nm.is_modification and all(name not in alphabet for name in nm.names):
                nm.modification_rate = 0.0
        self._reduce_nucleotide_list()

    def _reduce_nucleotide_list(self):
        kept = [nm for nm in self.masses if nm.mass == 0.0 or nm.modification_rate > 0.0]
        if len(kept) == len(self.masses):
            return
        self.masses = kept
        self._build_device_table()  # device rebuild; no host copy unless .table is read

    def print_masses(self):
        names = [n for nm in self.masses for n in nm.names]
        df = EXPLANATION_MASSES.sort("monoisotopic_mass")
        rep_col = df.get_column("nucleoside").to_list()
        keep = [i for i, rep in enumerate(rep_col) if rep in names]
        rates = [nm.modification_rate for nm in self.masses[1:]]
        cols = df.columns
        print(" | ".join(cols))
        for k, i in enumerate(keep):
            row = [df.get_column(c).to_list()[i] for c in cols]
            if k < len(rates):
                row[cols.index("modification_rate")] = rates[k]
            print(" | ".join(str(x) for x in row))

    def __repr__(self):
        return (f"DynamicProgrammingTable(rows={len(self.masses)}, compression_per_cell={self.compression_per_cell}, "
                f"precision={self.precision}, tolerance={self.tolerance}, seq={self.seq})")


def compute_sequence_length_bounds(dp_table: DynamicProgrammingTable) -> Tuple[int, int]:
    """(lower, upper) in one device walk (``sst_length_bounds``; the traversal does not depend on the direction)."""
    from .mass_explanation import _budget_int, _row_metadata

    max_modifications = round(dp_table.seq.modification_rate * dp_table.seq.max_len)
    target = int(round(dp_table.seq.su_mass / dp_table.precision, 0))
    threshold = int(np.ceil(dp_table.tolerance * dp_table.seq.obs_mass / dp_table.precision))
    dev = dp_table.device_table()
    _weights, is_mod, ind = _row_metadata(dp_table)
    cap = 0
    while True:
        try:
            return dev.ctx.length_bounds(dev, target, threshold, _budget_int(max_modifications), dp_table.seq.max_len, ind, is_mod, cap)
        except _cabi.MemoFull:
            cap = (cap or (1 << 16)) * 4
            if cap > (1 << 26):
                raise


def compute_sequence_length_bound(dp_table: DynamicProgrammingTable, dir: str) -> int:
    """Lower / upper bound on the number of nucleotides of any sequence explaining ``seq.su_mass`` (reference
    mass_table.py:343-487, including its first-visit memo semantics).  Runs on the device; both directions come
    out of the same walk, so the pair is kept for the second call a caller usually makes right away."""
    if dir not in ("lower", "upper"):
        raise NotImplementedError(f"Support for '{dir}' is currently not given.")
    rows = dp_table.masses
    key = (id(dp_table.device_table()), dp_table.seq.su_mass, dp_table.seq.obs_mass, dp_table.seq.max_len,
           dp_table.seq.modification_rate, dp_table.tolerance, tuple(m.modification_rate for m in rows))
    hit = getattr(dp_table, "_bounds_cache", None)
    if hit is None or hit[0] != key:
        hit = (key, compute_sequence_length_bounds(dp_table))
        dp_table._bounds_cache = hit
    return hit[1][0 if dir == "lower" else 1]
