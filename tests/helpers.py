"""Shared test helpers: golden loaders and alphabet builders (test infrastructure)."""
from __future__ import annotations

import gzip
import hashlib
import json
import pathlib

import numpy as np

from oracle import oracle_py as OP
from spectrseqtools_b200 import masses as M
from spectrseqtools_b200 import mass_table as MT

GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def load_json(name: str):
    path = GOLD / name
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as fh:
            return json.load(fh)
    return json.loads(path.read_text())


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(named) -> str:
    return OP.canonical_digest(named)


def full_weights():
    return [0] + sorted(M.EXPLANATION_MASSES.get_column("tolerated_integer_masses").to_list())


class FakeFrame:
    """Just enough of a frame for initialize_nucleotide_masses: three columns as lists."""

    def __init__(self, weights, names, rates):
        self._cols = {"tolerated_integer_masses": list(weights), "nucleoside": list(names), "modification_rate": list(rates)}

    def get_column(self, name):
        from spectrseqtools_b200._frame import Series

        return Series(name, self._cols[name])


def small_dp_table(weights, is_mod, rates, max_len, tolerance, device=None):
    """DynamicProgrammingTable over an arbitrary small alphabet (weights[0] == 0 is the implicit row)."""
    names = [("A" if not im else f"9x{w}") for w, im in zip(weights[1:], is_mod[1:])]
    frame = FakeFrame(weights[1:], names, rates[1:])
    seq = MT.SequenceInformation(max_len=max_len, su_mass=0.0, obs_mass=0.0, modification_rate=1.0)
    dp = MT.DynamicProgrammingTable(frame, 32, tolerance, 1e-3, seq, device=device)
    assert [m.mass for m in dp.masses] == list(weights), ([m.mass for m in dp.masses], weights)
    assert [m.is_modification for m in dp.masses] == [bool(x) for x in is_mod]
    return dp


def full_dp_table(max_len, tolerance=10e-6, modification_rate=0.5, su_mass=0.0, obs_mass=0.0):
    seq = MT.SequenceInformation(max_len=max_len, su_mass=su_mass, obs_mass=obs_mass, modification_rate=modification_rate)
    return MT.DynamicProgrammingTable(M.EXPLANATION_MASSES, 32, tolerance, M.TOLERANCE, seq)


def oracle_rows(dp):
    return [OP.Row(m.mass, m.is_modification, m.modification_rate) for m in dp.masses]
