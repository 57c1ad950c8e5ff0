#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (build container only) — golden vectors from the reference's own functions.

Runs the UNMODIFIED hot-path functions of /root/reference (loaded by oracle/ref_harness.py) on seeded
inputs and writes their outputs under tests/golden/.  The oracle (oracle_py.py / oracle.c) and the
CUDA path are both checked against these files; /root/reference itself is never needed at test time.

    python oracle/gen_golden.py            # everything except the 200 s full-table build
    python oracle/gen_golden.py --full     # also rebuild the full table with the reference loop

The full-alphabet table used for the explain cases is built with the C oracle and its SHA-256 is
asserted equal to the reference-built one recorded in tests/golden/tables_sha.json (``--full`` run).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import oracle_c as OC  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
from spectrseqtools_b200 import masses as M  # noqa: E402

GOLD = ROOT / "tests" / "golden"
TEST_SEQ = [("A",), ("A", "A"), ("G", "G"), ("C", "C"), ("U", "U"), ("C", "U", "A", "G"), ("C", "C", "U", "A", "G", "G")]
TOLS = [10e-6, 5e-6, 2e-6]
C2_MODS = "0C 0U 8U 2C 2U 9A 0A 04C 03U 01A 68A 7G 01G 071C 61A 62A 10G 51C 022G 2511U".split()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(named) -> str:
    if named is None:
        return "None"
    return hashlib.sha256("\n".join(",".join(t) for t in sorted(named)).encode()).hexdigest()[:16]


def full_rows():
    df = M.EXPLANATION_MASSES
    ims = df.get_column("tolerated_integer_masses").to_list()
    rates = df.get_column("modification_rate").to_list()
    order = sorted(range(len(ims)), key=lambda i: ims[i])
    rows = [(0, [], False, 0.0)]
    for i in order:
        names = M._INT_MASS_NAMES[ims[i]]
        rows.append((ims[i], names, M._INT_MASS_IS_MOD[ims[i]], rates[i]))
    return rows


def make_masses(ref, rows, universal_rate=None):
    out = []
    for mass, names, is_mod, rate in rows:
        if universal_rate is not None and is_mod and rate > universal_rate:
            rate = universal_rate
        out.append(ref.NucleotideMass(mass, list(names), is_mod, rate))
    return out


def seq_weight(seq):
    df = M.EXPLANATION_MASSES
    names = df.get_column("nucleoside").to_list()
    mono = df.get_column("monoisotopic_mass").to_list()
    return round(len(seq) * M.PHOSPHATE_LINK_MASS + sum(mono[names.index(x)] for x in seq), 5)


def gen_small_tables(ref):
    rng = np.random.default_rng(20260118)
    cases, arrays = [], {}
    for comp in (4, 8, 16, 32):
        made = 0
        while made < 10:
            k = int(rng.integers(1, 6))
            lo = 1 if made % 4 == 3 else comp
            w = [0] + sorted({int(x) for x in rng.integers(lo, 220, size=k)})
            mm = max(w) * int(rng.integers(2, 12)) + int(rng.integers(0, 6))
            try:
                t = ref.set_up_bit_table(w, mm, comp)
            except OverflowError:
                cases.append(dict(weights=w, max_mass=mm, compression=comp, raises="OverflowError"))
                continue
            key = f"t{len(arrays)}"
            arrays[key] = t
            cases.append(dict(weights=w, max_mass=mm, compression=comp, key=key, sha256=sha(t)))
            made += 1
    # narrow cells over weights the device build accepts (>= 32): the host re-pack of set_up_bit_table is checked on these
    rng2 = np.random.default_rng(20260122)
    for comp in (4, 8, 16):
        made = 0
        while made < 5:
            k = int(rng2.integers(1, 6))
            w = [0] + sorted({int(x) for x in rng2.integers(32, 260, size=k)})
            mm = max(w) * int(rng2.integers(2, 12)) + int(rng2.integers(0, 40))
            if made == 4:
                mm = (mm // comp + 1) * comp - 1  # (max_mass + 1) a multiple of the compression rate
            try:
                t = ref.set_up_bit_table(w, mm, comp)
            except OverflowError:
                cases.append(dict(weights=w, max_mass=mm, compression=comp, raises="OverflowError", narrow=True))
                made += 1
                continue
            key = f"t{len(arrays)}"
            arrays[key] = t
            cases.append(dict(weights=w, max_mass=mm, compression=comp, key=key, sha256=sha(t), narrow=True))
            made += 1
    # mask-wipe quirk: (35*w+1) % 32 == 0 for the maximum weight
    for w in ([0, 40, 53], [0, 33, 64 + 21]):
        mm = max(w) * 35
        t = ref.set_up_bit_table(w, mm, 32)
        key = f"t{len(arrays)}"
        arrays[key] = t
        cases.append(dict(weights=w, max_mass=mm, compression=32, key=key, sha256=sha(t), note="last-column quirk"))
    np.savez_compressed(GOLD / "tables_small.npz", **arrays)
    (GOLD / "tables_small.json").write_text(json.dumps(cases, indent=0) + "\n")
    print(f"small tables: {len(cases)} cases")


def gen_mass_tables(ref):
    """Byte-per-mass tables (set_up_mass_table, mass_table.py:292-316) for small alphabets, incl. a width whose
    (max_mass + 1) is a multiple of 32 (where the PACKED table's last-column mask wipes a whole word, the byte table
    keeps every cell)."""
    rng = np.random.default_rng(20260121)
    arrays, cases = {}, []
    for ci in range(12):
        k = int(rng.integers(1, 6))
        w = [0] + sorted({int(x) for x in rng.integers(32, 300, size=k)})
        mm = max(w) * int(rng.integers(2, 9)) + int(rng.integers(0, 40))
        if ci % 4 == 0:
            mm = (mm // 32 + 1) * 32 - 1
        t = ref.set_up_mass_table(w, mm)
        key = f"m{ci}"
        arrays[key] = t
        cases.append(dict(weights=w, max_mass=mm, key=key, sha256=sha(t)))
    np.savez_compressed(GOLD / "mass_tables_small.npz", **arrays)
    (GOLD / "mass_tables_small.json").write_text(json.dumps(cases) + "\n")
    print(f"byte tables: {len(cases)} cases")


def gen_table_shas(ref, full: bool):
    path = GOLD / "tables_sha.json"
    doc = json.loads(path.read_text()) if path.exists() else {}
    rows = full_rows()
    w_full = [r[0] for r in rows]
    acgu = [0] + sorted(r[0] for r in rows if r[1] and r[1][0] in ("A", "C", "G", "U") and not r[2])
    t0 = time.time()
    t = ref.set_up_bit_table(acgu, max(acgu) * 35, 32)
    doc["acgu"] = dict(weights=acgu, max_mass=max(acgu) * 35, compression=32, shape=list(t.shape), sha256=sha(t),
                       built_by="reference set_up_bit_table", seconds=round(time.time() - t0, 1))
    # 365045-max alphabet: wipes the whole last word of every row (numpy shift >= 64 -> 0)
    quirk = [w for w in w_full if w <= 365045]
    assert quirk[-1] == 365045
    t0 = time.time()
    t = ref.set_up_bit_table(quirk[:6] + [365045], 365045 * 35, 32)
    doc["quirk_365045"] = dict(weights=quirk[:6] + [365045], max_mass=365045 * 35, compression=32, shape=list(t.shape),
                               sha256=sha(t), last_word_all_zero=bool((t[:, -1] == 0).all()),
                               built_by="reference set_up_bit_table", seconds=round(time.time() - t0, 1))
    if full:
        t0 = time.time()
        t = ref.set_up_bit_table(w_full, max(w_full) * 35, 32)
        doc["full"] = dict(weights=w_full, max_mass=max(w_full) * 35, compression=32, shape=list(t.shape), sha256=sha(t),
                           built_by="reference set_up_bit_table", seconds=round(time.time() - t0, 1))
    elif pathlib.Path("/tmp/ref_full_table.npy").exists():
        t = np.load("/tmp/ref_full_table.npy")
        doc["full"] = dict(weights=w_full, max_mass=max(w_full) * 35, compression=32, shape=list(t.shape), sha256=sha(t),
                           built_by="reference set_up_bit_table (cached /tmp/ref_full_table.npy)")
    path.write_text(json.dumps(doc, indent=1) + "\n")
    print("table shas:", {k: v["sha256"][:12] for k, v in doc.items()})
    return doc


def gen_explain_small(ref):
    """Random small alphabets where the Python reference is fast: explain (memo / no memo), validity, length bounds."""
    rng = np.random.default_rng(20260119)
    cases = []
    for ci in range(400):
        k = int(rng.integers(2, 7))
        weights = sorted({int(x) for x in rng.integers(32, 400, size=k)})
        rows = [(0, [], False, 0.0)]
        for w in weights:
            is_mod = bool(rng.random() < 0.6)
            rate = float(rng.choice([0.0, 0.1, 0.25, 0.5, 1.0])) if is_mod else 1.0
            rows.append((w, [f"n{w}"], is_mod, rate))
        rows = [r for r in rows if r[0] == 0 or r[3] > 0.0]
        if len(rows) < 2:
            continue
        w_list = [r[0] for r in rows]
        max_mass = max(w_list) * 35
        table = ref.set_up_bit_table(w_list, max_mass, 32)
        masses = make_masses(ref, rows)
        max_len = int(rng.integers(1, 14))
        seq = ref.SequenceInformation(max_len=max_len, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
        dp = H.DuckTable(table, masses, seq, precision=1e-3, tolerance=float(rng.choice([1e-3, 5e-3, 2e-2])))
        mass_names = {r[0]: list(r[1]) for r in rows if r[0]}
        ref.MASS_NAMES.clear(); ref.MASS_NAMES.update(mass_names)  # noqa: E702
        for _ in range(4):
            n_nt = int(rng.integers(0, 9))
            true = int(sum(rng.choice(w_list[1:], size=n_nt))) if n_nt else 0
            mass = (true + int(rng.integers(-3, 4))) * 1e-3
            threshold = None if rng.random() < 0.4 else float(rng.integers(0, 12)) * 1e-3
            mm = math.inf if rng.random() < 0.3 else int(rng.integers(0, 6))
            entry = dict(weights=w_list, is_mod=[bool(r[2]) for r in rows], rates=[r[3] for r in rows],
                         max_len=max_len, tolerance=dp.tolerance, mass=mass, threshold=threshold,
                         max_modifications=None if mm == math.inf else mm)
            for memo in (True, False):
                captured = {}
                real_convert = ref.convert_nucleotide_masses_to_names

                def spy(solutions, _c=captured):
                    _c["solutions"] = [list(s) for s in solutions]
                    return real_convert(solutions=solutions)

                ref.explain_mass_with_table.__globals__["convert_nucleotide_masses_to_names"] = spy
                try:
                    res = ref.explain_mass_with_table(mass, dp, max_modifications=mm, threshold=threshold, with_memo=memo)
                    tag = "memo" if memo else "nomemo"
                    entry[f"solutions_{tag}"] = captured["solutions"]
                    entry[f"named_{tag}"] = None if res.explanations is None else sorted(list(t) for t in res.explanations)
                except NotImplementedError:
                    entry[f"solutions_{'memo' if memo else 'nomemo'}"] = "NotImplementedError"
                finally:
                    ref.explain_mass_with_table.__globals__["convert_nucleotide_masses_to_names"] = real_convert
            try:
                entry["is_valid"] = bool(ref.is_valid_mass(mass, dp, threshold))
            except NotImplementedError:
                entry["is_valid"] = "NotImplementedError"
            if mass > 0:
                dp.seq = ref.SequenceInformation(max_len=max_len, su_mass=mass, obs_mass=mass * 1.1, modification_rate=0.5)
                for d in ("lower", "upper"):
                    try:
                        entry[f"bound_{d}"] = int(ref.compute_sequence_length_bound(dp, d))
                    except NotImplementedError:
                        entry[f"bound_{d}"] = "NotImplementedError"
                entry["bound_obs_mass"] = mass * 1.1
            cases.append(entry)
    import gzip

    with gzip.open(GOLD / "explain_small.json.gz", "wt", compresslevel=9) as fh:
        json.dump(cases, fh)
    n_sol = sum(len(c["solutions_memo"]) for c in cases if isinstance(c["solutions_memo"], list))
    differ = sum(1 for c in cases if c["solutions_memo"] != c["solutions_nomemo"])
    print(f"explain_small: {len(cases)} cases, {n_sol} memo solutions, {differ} cases where memo != no-memo")


def gen_explain_full(ref, table_sha: str):
    rows = full_rows()
    w_full = [r[0] for r in rows]
    table = OC.build_bit_table(w_full, max(w_full) * 35, 32)
    assert sha(table) == table_sha, "C-oracle table differs from the reference-built table"
    ref.MASS_NAMES.clear(); ref.MASS_NAMES.update(M._INT_MASS_NAMES)  # noqa: E702
    ref.IS_MOD.clear(); ref.IS_MOD.update(M._INT_MASS_IS_MOD)  # noqa: E702
    min_w = w_full[1]
    out = dict(table_sha256=table_sha, unit_test_cases=[], random_cases=[], validity_cases=[])
    # --- the reference's own unit-test inputs (tests/test_explain_masses.py:34-66,97-130)
    for seq in TEST_SEQ:
        mass = seq_weight(seq)
        max_len = int(mass / M.TOLERANCE / min_w)
        for tol in TOLS:
            masses = make_masses(ref, rows, universal_rate=0.5)
            sinfo = ref.SequenceInformation(max_len=max_len, su_mass=mass, obs_mass=mass, modification_rate=0.5)
            dp = H.DuckTable(table, masses, sinfo, precision=M.TOLERANCE, tolerance=tol)
            mm = round(0.5 * len(seq))
            entry = dict(seq=list(seq), mass=mass, tolerance=tol, max_len=max_len, max_modifications=mm)
            for memo in (True, False):
                t0 = time.time()
                res = ref.explain_mass_with_table(mass, dp, max_modifications=mm, with_memo=memo).explanations
                tag = "memo" if memo else "nomemo"
                entry[f"n_{tag}"] = None if res is None else len(res)
                entry[f"digest_{tag}"] = digest(res)
                entry[f"seconds_{tag}"] = round(time.time() - t0, 4)
                if res is not None and len(res) <= 16:
                    entry[f"set_{tag}"] = sorted(list(t) for t in res)
            t0 = time.time()
            rec = ref.explain_mass_with_recursion(mass, dp, max_modifications=mm).explanations
            entry["n_recursion"] = None if rec is None else len(rec)
            entry["digest_recursion"] = digest(rec)
            entry["seconds_recursion"] = round(time.time() - t0, 4)
            out["unit_test_cases"].append(entry)
            print("  unit", "".join(seq), tol, entry["n_memo"], entry["n_nomemo"], entry["n_recursion"])
    # --- production-like random differences (budgets cannot bind) and budget-bound variants
    rng = np.random.default_rng(20260120)
    for ci in range(60):
        n_nt = int(rng.integers(1, 4))
        picks = rng.choice(len(w_full) - 1, size=n_nt) + 1
        true = int(sum(w_full[p] for p in picks))
        mass = true * 1e-3 * (1 + float(rng.uniform(-3e-6, 3e-6)))
        thr = float(rng.choice([0.02, 0.06, 0.13, 0.26]))
        bound = ci % 3 == 2
        max_len = int(rng.integers(1, 5)) if bound else int(rng.integers(20, 60))
        mm = int(rng.integers(0, 3)) if bound else round(0.5 * max_len)
        masses = make_masses(ref, rows, universal_rate=0.5)
        sinfo = ref.SequenceInformation(max_len=max_len, su_mass=mass, obs_mass=mass, modification_rate=0.5)
        dp = H.DuckTable(table, masses, sinfo, precision=M.TOLERANCE, tolerance=10e-6)
        entry = dict(mass=mass, threshold=thr, max_len=max_len, max_modifications=mm, n_nt=n_nt, budget_bound=bound)
        for memo in (True, False):
            res = ref.explain_mass_with_table(mass, dp, max_modifications=mm, threshold=thr, with_memo=memo).explanations
            tag = "memo" if memo else "nomemo"
            entry[f"n_{tag}"] = None if res is None else len(res)
            entry[f"digest_{tag}"] = digest(res)
        out["random_cases"].append(entry)
    # --- validity probes incl. the table edge
    limit = table.shape[1] * 32
    masses = make_masses(ref, rows, universal_rate=0.5)
    sinfo = ref.SequenceInformation(max_len=35, su_mass=0.0, obs_mass=0.0, modification_rate=0.5)
    dp = H.DuckTable(table, masses, sinfo, precision=M.TOLERANCE, tolerance=10e-6)
    probes = [0.0, 0.0005, 0.3, 305.0, 305.042, 305.0425, 610.084, 611.07, 900.0, 2982.949, 2982.95, 5000.1234,
              (limit - 1) * 1e-3, (limit - 300) * 1e-3, limit * 1e-3, (max(w_full) * 35) * 1e-3, (max(w_full) * 35 + 1) * 1e-3]
    probes += [float(x) for x in rng.uniform(300, 3200, size=80)]
    for mass in probes:
        for thr in (None, 0.0, 0.004, 0.06):
            try:
                v = bool(ref.is_valid_mass(mass, dp, thr))
            except NotImplementedError:
                v = "NotImplementedError"
            out["validity_cases"].append(dict(mass=mass, threshold=thr, valid=v))
    (GOLD / "explain_full.json").write_text(json.dumps(out, indent=0) + "\n")
    print(f"explain_full: {len(out['unit_test_cases'])} unit, {len(out['random_cases'])} random, {len(out['validity_cases'])} validity")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="rebuild the full table with the reference loop (~200 s)")
    ap.add_argument("--only", default="", help="comma list of: small_tables,mass_tables,shas,explain_small,explain_full")
    args = ap.parse_args()
    if not H.available():
        raise SystemExit("/root/reference not present: golden vectors can only be regenerated in the build container")
    GOLD.mkdir(parents=True, exist_ok=True)
    ref = H.load_reference(dict(M._INT_MASS_NAMES), dict(M._INT_MASS_IS_MOD))
    only = set(filter(None, args.only.split(",")))
    if not only or "small_tables" in only:
        gen_small_tables(ref)
    if not only or "mass_tables" in only:
        gen_mass_tables(ref)
    doc = None
    if not only or "shas" in only:
        doc = gen_table_shas(ref, args.full)
    if not only or "explain_small" in only:
        gen_explain_small(ref)
    if not only or "explain_full" in only:
        doc = doc or json.loads((GOLD / "tables_sha.json").read_text())
        if "full" not in doc:
            raise SystemExit("tables_sha.json has no 'full' entry yet: run with --full once")
        gen_explain_full(ref, doc["full"]["sha256"])


if __name__ == "__main__":
    main()
