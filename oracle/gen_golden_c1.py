#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (build container only) — golden vectors for BASELINE config C1 and the rows N3 / N4.

C1 = the reference's own fixture ``tests/testcases/test_02`` (105 simulated fragments of GAUUGUCGUG + meta; copied
as data under tests/golden/c1/) pushed through the reference's UNMODIFIED function bodies, loaded by
oracle/ref_harness.py (nothing of the reference's code is copied):

  * ``classify_fragments`` (fragment_classification.py:17-101) -> the classified frame,
  * ``Predictor.filter_by_explanation`` (prediction.py:170-202) with ``collect_diff_explanations_for_su`` (:261-284),
    ``collect_explanations_per_side`` (:286-329), ``_reduce_alphabet`` (:204-227), ``calculate_explanations``
    (common.py:47-65) and the table's ``adapt_individual_modification_rates_by_alphabet_reduction`` /
    ``_reduce_nucleotide_list`` (mass_table.py:94-121, which REBUILDS the table with ``set_up_bit_table``),

for two alphabets: ``acgu`` (C1 proper: every modification at rate 0 -> 5-row table, as cli.py:112-139 does when only
the four bases are seen as singletons) and ``mods12`` (A/C/G/U + 8 modifications, so that the fixed point has
something to remove).  The frames are shaped by the repo's polars stand-in (polars is not in the image); every number
comes out of the reference's own arithmetic.  A spy around ``calculate_explanations`` / ``is_valid_mass`` logs every
call in order, which pins the pair generator (N3) call for call.

Output: tests/golden/c1.json
"""
from __future__ import annotations

import json
import pathlib
import sys
import time

import yaml

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_harness as H  # noqa: E402
from spectrseqtools_b200 import _frame as pl  # noqa: E402
from spectrseqtools_b200 import masses as M  # noqa: E402

FIX = ROOT / "tests" / "golden" / "c1"
MODS12 = ["A", "C", "G", "U", "0A", "9A", "0C", "0U", "8U", "2C", "7G", "01G"]


def alphabet_rows(keep):
    """(mass, names, is_mod, rate) per table row for the representatives in `keep`, leading 0 row included."""
    df = M.EXPLANATION_MASSES
    rows = [(0, [], False, 0.0)]
    for m, n, r in sorted(zip(df.get_column("tolerated_integer_masses").to_list(), df.get_column("nucleoside").to_list(),
                              df.get_column("modification_rate").to_list())):
        if n in keep:
            rows.append((int(m), list(M._INT_MASS_NAMES[m]), bool(M._INT_MASS_IS_MOD[m]), 1.0 if n in M.UNMODIFIED_BASES else float(r)))
    return rows


def make_table(ref, rows, seq, tolerance):
    """The reference's DynamicProgrammingTable state, built by its own methods (no polars in them)."""
    class Table(ref.RefTableMethods):
        def print_masses(self):  # printing only (needs polars upstream)
            pass

    dp = Table()
    dp.compression_per_cell, dp.tolerance, dp.precision, dp.seq = 32, tolerance, M.TOLERANCE, seq
    dp.masses = [ref.NucleotideMass(m, list(n), im, r) for m, n, im, r in rows]
    dp.table = None
    dp._adapt_individual_modification_rates_by_universal_one()  # mass_table.py:75
    if dp.table is None:                                        # :78-84 (load_dp_table = the same builder)
        w = [x.mass for x in dp.masses]
        dp.table = ref.set_up_bit_table(w, max(w) * 35, 32)
    return dp


def run(ref, name, keep):
    meta = yaml.safe_load((FIX / "fragments.meta.yaml").read_text())
    frags = pl.read_csv(FIX / "fragments.tsv", separator="\t")
    frags = frags.with_columns(pl.col("observed_mass").alias("observed_mass"), pl.col("true_mass_with_backbone").alias("true_mass"))
    breakage = M.build_breakage_dict(mass_5_prime=meta["label_mass_5T"], mass_3_prime=meta["label_mass_3T"])
    seq_obs = meta["sequence_mass"]
    seq_su = seq_obs - [w * M.TOLERANCE for w in breakage if "START_END" in breakage[w]][0]   # tests/test_prediction.py:145-154
    rows = alphabet_rows(keep)
    min_w = min(r[0] for r in rows if r[0])
    seq = ref.SequenceInformation(max_len=int(seq_su / M.TOLERANCE / min_w), su_mass=seq_su, obs_mass=seq_obs, modification_rate=0.5)
    t0 = time.time()
    dp = make_table(ref, rows, seq, M.MATCHING_THRESHOLD)
    log = {"is_valid": [], "explain": []}
    states = {}  # rows in the table at call time -> (weights, rates): the alphabet only ever shrinks

    real_valid, real_calc = ref.is_valid_mass, ref.calculate_explanations

    def spy_valid(mass, dp_table, threshold=None):
        out = real_valid(mass=mass, dp_table=dp_table, threshold=threshold)
        states.setdefault(len(dp_table.masses), [[m.mass for m in dp_table.masses], [m.modification_rate for m in dp_table.masses]])
        log["is_valid"].append([mass, threshold, bool(out), len(dp_table.masses)])
        return out

    def spy_calc(diff, threshold, dp_table):
        out = real_calc(diff=diff, threshold=threshold, dp_table=dp_table)
        states.setdefault(len(dp_table.masses), [[m.mass for m in dp_table.masses], [m.modification_rate for m in dp_table.masses]])
        log["explain"].append([diff, threshold, None if out is None else sorted(list(e.nucleosides) for e in out), len(dp_table.masses)])
        return out

    g = ref.classify_fragments.__globals__
    g["is_valid_mass"], g["calculate_explanations"], g["PHOSPHATE_LINK_MASS"] = spy_valid, spy_calc, M.PHOSPHATE_LINK_MASS
    try:
        classified = ref.classify_fragments(fragment_masses=frags, dp_table=dp, breakage_dict=breakage,
                                            intensity_cutoff=M.DEFAULT_INTENSITY_CUTOFF)
        n_valid_calls = len(log["is_valid"])
        classify_log = log["is_valid"]
        log["is_valid"] = []
        # Predictor.predict prologue (prediction.py:68-80), then the fixed point
        cur = classified.with_row_index(name="orig_index").sort("standard_unit_mass").with_row_index(name="index")

        class P(ref.RefPredictor):
            pass

        pr = P()
        pr.dp_table, pr.explanation_masses = dp, M.EXPLANATION_MASSES.filter([n in keep for n in M.EXPLANATION_MASSES.get_column("nucleoside").to_list()])
        rounds_before = len(log["explain"])
        final, explanations = pr.filter_by_explanation(cur)
    finally:
        g["is_valid_mass"], g["calculate_explanations"] = real_valid, real_calc
    out = {
        "name": name, "alphabet": keep, "tolerance": M.MATCHING_THRESHOLD, "breakage": {str(k): v for k, v in breakage.items()},
        "seq": {"max_len": seq.max_len, "su_mass": seq.su_mass, "obs_mass": seq.obs_mass, "modification_rate": 0.5},
        "start_weights": [r[0] for r in rows if r[0] == 0 or r[3] > 0],
        "table_states": {str(k): v for k, v in states.items()},
        "classify_calls": classify_log,
        "classified": {c: classified.get_column(c).to_list() for c in ("fragment_index", "observed_mass", "standard_unit_mass", "breakage", "is_singleton")},
        "explain_calls": log["explain"][rounds_before:],
        "revalidate_calls": log["is_valid"],
        "final_orig_index": final.get_column("orig_index").to_list(),
        "final_weights": [m.mass for m in dp.masses],
        "final_names": [list(m.names) for m in dp.masses],
        "explanations": [[k, None if v is None else sorted(list(e.nucleosides) for e in v)] for k, v in explanations.items()],
    }
    print(f"{name}: {n_valid_calls} validity + {len(out['explain_calls'])} explanation calls, classified {len(classified)} rows, "
          f"final {len(final)} rows, alphabet {len(rows)} -> {len(dp.masses)} rows, {time.time() - t0:.1f} s")
    return out


def main():
    if not H.available():
        raise SystemExit("reference sources not found (/root/reference or baseline/_ref)")
    ref = H.load_reference(dict(M._INT_MASS_NAMES), dict(M._INT_MASS_IS_MOD), pl=pl)
    doc = [run(ref, "acgu", ["A", "C", "G", "U"]), run(ref, "mods12", MODS12)]
    (ROOT / "tests" / "golden" / "c1.json").write_text(json.dumps(doc))


if __name__ == "__main__":
    main()
