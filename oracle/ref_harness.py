"""TEST INFRASTRUCTURE (build container only) — run the reference's own hot-path functions, polars-free.

The reference package cannot be imported here (``import polars`` fails at masses.py:2), but the bodies
of its hot-path functions only use numpy + itertools.  This module reads the UNMODIFIED source files
under /root/reference at run time, pulls the named top-level definitions out with ``ast`` and executes
them in a namespace that supplies the alphabet dictionaries.  Nothing is copied into the repo; the
only products are the golden vectors written by ``oracle/gen_golden.py``.

/root/reference does not exist on the GPU box: nothing under tests/ -m gpu, smoke() or bench.py may
import this file.
"""
from __future__ import annotations

import ast
import itertools
import pathlib
import sys
from dataclasses import dataclass
from typing import List, Set, Tuple

import numpy as np

REF_ROOT = pathlib.Path("/root/reference/spectrseqtools")

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
    "fragment_classification.py": ["is_singleton"],
}


def available() -> bool:
    return all((REF_ROOT / f).is_file() for f in _WANTED)


def load_reference(mass_names: dict, is_mod: dict):
    """Return a namespace object holding the reference's functions bound to the given alphabet maps."""
    ns: dict = {
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
