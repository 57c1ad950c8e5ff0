"""TEST INFRASTRUCTURE — run the reference's own hot-path functions, polars-free.

The reference package cannot be imported (``import polars`` fails at masses.py:2 in this image), but the
bodies of its hot-path functions only use numpy + itertools.  This module reads the UNMODIFIED source files of
the reference at run time, pulls the named definitions out with ``ast`` and executes them in a namespace that
supplies the alphabet dictionaries.  Nothing is copied into the git history: the sources are read from
/root/reference (build container) or from ``baseline/_ref/`` — the git-ignored copy of the reference that
``__graft_entry__.build()`` makes in the build container and that travels to the GPU box with the snapshot
(the prescribed place for the reference install of the bench's reference arm).

Users: ``oracle/gen_golden*.py`` (golden vectors), ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference``
(the reference's own functions timed on the host cores), and tests.  Never the product.
"""
from __future__ import annotations

import ast
import itertools
import pathlib
import sys
from dataclasses import dataclass
from typing import List, Set, Tuple

import numpy as np

_REPO = pathlib.Path(__file__).resolve().parents[1]


def _find_root() -> pathlib.Path:
    for cand in (pathlib.Path("/root/reference"), _REPO / "baseline" / "_ref"):
        if (cand / "spectrseqtools" / "mass_explanation.py").is_file():
            return cand
    return pathlib.Path("/root/reference")


REF_TOP = _find_root()                 # holds spectrseqtools/ and tests/
REF_ROOT = REF_TOP / "spectrseqtools"

_WANTED = {
    "mass_table.py": [
        "SequenceInformation",
        "NucleotideMass",
        "set_up_bit_table",
        "select_table_building_settings",
        "set_up_mass_table",
        "compute_sequence_length_bound",
    ],
    "mass_explanation.py": [
        "MassExplanations",
        "is_valid_mass",
        "explain_mass_with_table",
        "explain_mass_with_recursion",
        "convert_nucleotide_masses_to_names",
    ],
    "fragment_classification.py": ["is_singleton", "filter_by_sequence_mass", "classify_fragments"],
    "common.py": ["Explanation", "calculate_error_threshold", "calculate_explanations"],
}
# methods pulled out of class bodies (the classes themselves need polars / PuLP at definition time)
_WANTED_METHODS = {
    "prediction.py": {"Predictor": ["filter_by_explanation", "_reduce_alphabet", "collect_diff_explanations_for_su",
                                    "collect_explanations_per_side"]},
    "mass_table.py": {"DynamicProgrammingTable": ["_adapt_individual_modification_rates_by_universal_one",
                                                  "adapt_individual_modification_rates_by_alphabet_reduction",
                                                  "_reduce_nucleotide_list"]},
}


def available() -> bool:
    return all((REF_ROOT / f).is_file() for f in _WANTED)


def load_reference(mass_names: dict, is_mod: dict, pl=None):
    """Return a namespace object holding the reference's functions bound to the given alphabet maps.

    ``pl``: module standing in for polars in the frame-shaped functions (``classify_fragments`` and the ``Predictor``
    methods); defaults to the repo's minimal stand-in (``spectrseqtools_b200/_frame.py``) when polars is missing.
    The extracted ``Predictor`` / ``DynamicProgrammingTable`` methods are exposed as ``RefPredictor`` /
    ``RefTableMethods``: plain classes holding the unmodified method bodies."""
    if pl is None:
        try:
            import polars as pl  # noqa: F811
        except ImportError:
            sys.path.insert(0, str(_REPO))
            from spectrseqtools_b200 import _frame as pl  # noqa: F811
    ns: dict = {
        "pl": pl,
        "ERROR_METHOD": "l1_norm",
        "MAX_VARIANCE": 1,
        "Optional": __import__("typing").Optional,
        "np": np,
        "dataclass": dataclass,
        "List": List,
        "Set": Set,
        "Tuple": Tuple,
        "product": itertools.product,
        "combinations_with_replacement": itertools.combinations_with_replacement,
        "chain": itertools.chain,
        "MASS_NAMES": mass_names,
        "IS_MOD": is_mod,
        "DynamicProgrammingTable": object,  # annotation only
        "MAX_SEQ_LENGTH": 35,
    }
    for fname, wanted in _WANTED.items():
        src = (REF_ROOT / fname).read_text()
        tree = ast.parse(src)
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in wanted:
                code = compile(ast.Module(body=[node], type_ignores=[]), str(REF_ROOT / fname), "exec")
                exec(code, ns)
    missing = [w for ws in _WANTED.values() for w in ws if w not in ns]
    if missing:
        raise RuntimeError(f"reference definitions not found: {missing}")
    for fname, classes in _WANTED_METHODS.items():
        tree = ast.parse((REF_ROOT / fname).read_text())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in classes:
                body = [m for m in node.body if isinstance(m, ast.FunctionDef) and m.name in classes[node.name]]
                found = {m.name for m in body}
                if found != set(classes[node.name]):
                    raise RuntimeError(f"{fname}:{node.name}: methods not found: {set(classes[node.name]) - found}")
                holder = ast.ClassDef(name="Ref" + ("Predictor" if node.name == "Predictor" else "TableMethods"), bases=[], keywords=[],
                                      body=body, decorator_list=[], type_params=[])
                mod = ast.Module(body=[holder], type_ignores=[])
                ast.fix_missing_locations(mod)
                exec(compile(mod, str(REF_ROOT / fname), "exec"), ns)
    ns.setdefault("PHOSPHATE_LINK_MASS", None)  # bound by the caller (masses.PHOSPHATE_LINK_MASS)

    class _NS:
        pass

    out = _NS()
    for k, v in ns.items():
        setattr(out, k, v)
    return out


class DuckTable:
    """Stand-in for the reference's DynamicProgrammingTable holder (mass_table.py:52-59 attributes)."""

    def __init__(self, table, masses, seq, precision=1e-3, tolerance=10e-6, compression_per_cell=32):
        self.table = table
        self.masses = masses
        self.seq = seq
        self.precision = precision
        self.tolerance = tolerance
        self.compression_per_cell = compression_per_cell


if __name__ == "__main__":
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
    from spectrseqtools_b200 import masses as M

    ref = load_reference(M._INT_MASS_NAMES, M._INT_MASS_IS_MOD)
    print("loaded:", [k for ws in _WANTED.values() for k in ws])
