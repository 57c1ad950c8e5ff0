"""Alphabet and constants of the mass-explanation path (host side, polars optional).

Mirrors the public names of the reference's ``spectrseqtools/masses.py`` (constants :10-34, element
table :38-50, ``initialize_nucleotide_df`` :53-88, ``EXPLANATION_MASSES`` :91, ``NUC_REPS`` :94-102,
``build_breakage_dict`` :110-160) so downstream code imports them from the same place.  The data come
from ``assets/alphabet.json`` (imported once by ``tools/import_alphabet.py``).

The grouping itself is done in plain Python so that it does not depend on polars being installed; the
result is wrapped in a real ``polars.DataFrame`` when polars is importable and in the small stand-in
from ``_frame.py`` otherwise.
"""
from __future__ import annotations

import json
import math
from itertools import product
from pathlib import Path

try:  # pragma: no cover - depends on the image
    import polars as _pl

    if getattr(_pl, "__spectrseq_shim__", False):
        raise ImportError
    _HAVE_POLARS = True
except ImportError:  # polars absent: use the stand-in
    from . import _frame as _pl

    _HAVE_POLARS = False

_COLS = ["nucleoside", "canonical_name", "monoisotopic_mass", "modification_rate"]

# TODO upstream: RNA only
UNMODIFIED_BASES = ["A", "C", "G", "U"]

DEFAULT_INTENSITY_CUTOFF = 115000

# c/y-only breakage dictionary unless switched on
FULL_BREAKAGE_DICT = False

# integer masses packed per table cell (2 bits each)
COMPRESSION_RATE = 32

DECIMAL_PLACES = 3
TOLERANCE = 10 ** (-DECIMAL_PLACES)

# relative matching threshold (10 ppm)
MATCHING_THRESHOLD = 10e-6

_ASSET = Path(__file__).resolve().parent / "assets" / "alphabet.json"
with open(_ASSET, encoding="utf-8") as _fh:
    _DOC = json.load(_fh)
assert _DOC["columns"] == _COLS

ELEMENT_MASSES = dict(_DOC["elements"])

# phosphate link between two nucleosides: P + 2 O - H+
PHOSPHATE_LINK_MASS = ELEMENT_MASSES["P"] + 2 * ELEMENT_MASSES["O"] - ELEMENT_MASSES["H+"]


def _round_like_polars(value: float, decimals: int) -> float:
    """Scale, round to the nearest integer, unscale (what ``pl.Expr.round`` does for f64).

    NOT ``round(value, decimals)``: CPython rounds the exact binary value, which turns
    A = 267.09675 into 267.0967 where polars (and numpy) give 267.0968.
    """
    scale = 10.0**decimals
    return round(value * scale) / scale


def _grouped_alphabet():
    """Rows of EXPLANATION_MASSES as plain Python tuples (reference masses.py:53-88)."""
    groups = {}  # rounded mass -> [representative, members, max rate]; dicts keep first-seen order
    for nucleoside, _canonical, mass, rate in _DOC["nucleosides"]:
        key = _round_like_polars(mass, DECIMAL_PLACES + 1)
        g = groups.get(key)
        if g is None:
            groups[key] = [nucleoside, [nucleoside], rate]
        else:
            if nucleoside not in g[1]:
                g[1].append(nucleoside)
            g[2] = max(g[2], rate)
    rows = []
    for mass, (rep, members, rate) in groups.items():
        mz = mass + (PHOSPHATE_LINK_MASS - ELEMENT_MASSES["H+"])
        integer_mass = int(round((mass + PHOSPHATE_LINK_MASS) / TOLERANCE))
        rows.append((mass, rep, members, rate, mz, integer_mass))
    return rows


def initialize_nucleotide_df():
    rows = _grouped_alphabet()
    names = [
        "monoisotopic_mass",
        "nucleoside",
        "nucleoside_list",
        "modification_rate",
        "theoretical_mz",
        "tolerated_integer_masses",
    ]
    return _pl.DataFrame({n: [r[j] for r in rows] for j, n in enumerate(names)})


EXPLANATION_MASSES = initialize_nucleotide_df()

# nucleoside -> representative of its equal-mass group
NUC_REPS = {nuc: rep for _m, rep, members, *_ in _grouped_alphabet() for nuc in members}

# integer mass -> representatives / "is a modification" (reference mass_explanation.py:17-42).
# Kept here (host data) so the explanation module does not need frame joins.
_INT_MASS_NAMES: dict = {}
for _m, _rep, _members, _rate, _mz, _im in _grouped_alphabet():
    _INT_MASS_NAMES.setdefault(_im, []).append(_rep)
_INT_MASS_IS_MOD = {im: any(n not in UNMODIFIED_BASES for n in names) for im, names in _INT_MASS_NAMES.items()}


def build_breakage_dict(mass_5_prime, mass_3_prime):
    """Integer (mDa, truncated) mass offsets of every 5'/3' end-type pair -> list of "<start>_<end>" labels."""
    em = ELEMENT_MASSES
    start = {
        "START": mass_5_prime - em["O"] - em["H+"],  # tag replaces O, no H
        "c/y": em["H+"],
    }
    end = {
        "END": mass_3_prime - em["P"] - 3 * em["O"] - 2 * em["H+"],  # tag replaces PO3H
        "c/y": -em["H+"],
    }
    if FULL_BREAKAGE_DICT:
        po3h2 = em["P"] + 3 * em["O"] + 2 * em["H+"]
        p2o = em["P"] + 2 * em["O"]
        oh = em["O"] + em["H+"]
        start.update({"a/w": po3h2, "b/x": p2o, "d/z": -oh})
        end.update({"a/w": -po3h2, "b/x": -p2o, "d/z": oh})

    out = {}
    for s, e in product(start, end):
        out.setdefault(int((start[s] + end[e]) / TOLERANCE), []).append(f"{s}_{e}")
    return out
