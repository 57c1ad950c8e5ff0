#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (build container only) — golden vectors for the fused fragment classification (N2).

Runs the reference's UNMODIFIED ``is_valid_mass`` (mass_explanation.py:45-89) and ``is_singleton``
(fragment_classification.py:104-119), loaded by oracle/ref_harness.py, on every (fragment x breakage) pair of a
seeded synthetic ladder, with the standard-unit mass and threshold computed by the same float expressions as
fragment_classification.py:39-60.  Output: tests/golden/classify.json.
"""
from __future__ import annotations

import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import oracle_c as OC  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
from spectrseqtools_b200 import masses as M  # noqa: E402


def main():
    ref = H.load_reference(M._INT_MASS_NAMES, M._INT_MASS_IS_MOD)
    full = sorted(M.EXPLANATION_MASSES.get_column("tolerated_integer_masses").to_list())
    acgu = sorted(int(m) for m, names in M._INT_MASS_NAMES.items() if any(n in M.UNMODIFIED_BASES for n in names))
    cases = []
    for name, weights, full_dict, tol, seed in (("acgu", acgu, False, 10e-6, 1), ("full", full, True, 10e-6, 2), ("full_5ppm", full, False, 5e-6, 3)):
        w = [0] + list(weights)
        table = OC.build_bit_table(w, max(w) * 35, 32)
        masses = [ref.NucleotideMass(m, [], False, 1.0) for m in w]
        seq = ref.SequenceInformation(max_len=40, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
        dp = H.DuckTable(table, masses, seq, precision=1e-3, tolerance=tol)
        old = M.FULL_BREAKAGE_DICT
        try:
            M.FULL_BREAKAGE_DICT = full_dict
            breakage = M.build_breakage_dict(555.1294, 455.1491)
        finally:
            M.FULL_BREAKAGE_DICT = old
        rng = np.random.default_rng(20260118 + seed)
        obs = []
        offs = sorted(breakage)
        for _ in range(6):  # ladders of a random oligo, observed under a random breakage, plus noise and decoys
            L = int(rng.integers(3, 30))
            pick = rng.choice(weights, size=L)
            su = np.cumsum(pick) * 1e-3
            off = offs[int(rng.integers(len(offs)))] * 1e-3
            obs.extend(((su + off) * (1 + rng.uniform(-tol / 2, tol / 2, size=L))).tolist())
        obs.extend(rng.uniform(200.0, 9000.0, size=20).tolist())
        obs.extend([0.0, 0.3, 305.04, 22160.9 + 0.912, 23000.0])  # tiny, near the table end, beyond it
        valid, single = [], []
        for bw in breakage:
            for x in obs:
                su_mass = x - (bw * dp.precision)
                thr = dp.tolerance * x
                try:
                    v = 1 if ref.is_valid_mass(mass=su_mass, dp_table=dp, threshold=thr) else 0
                except NotImplementedError:
                    v = 2
                valid.append(v)
                single.append(1 if ref.is_singleton(mass=su_mass, integer_masses=[m.mass for m in dp.masses], dp_table=dp, threshold=thr) else 0)
        cases.append({"name": name, "weights": w, "tolerance": tol, "breakage": {str(k): v for k, v in breakage.items()},
                      "observed": obs, "valid": valid, "singleton": single})
        print(name, "pairs", len(valid), "valid", sum(1 for v in valid if v == 1), "oot", sum(1 for v in valid if v == 2), "singleton", sum(single))
    (ROOT / "tests" / "golden" / "classify.json").write_text(json.dumps(cases))


if __name__ == "__main__":
    main()
