"""CPU: the oracle (Python + C restatements) against golden vectors made by the reference's own functions."""
import math

import numpy as np
import pytest

import helpers as Hh
from oracle import oracle_c as OC
from oracle import oracle_py as OP
from spectrseqtools_b200 import masses as M


def test_small_tables_match_reference():
    cases = Hh.load_json("tables_small.json")
    arrays = np.load(Hh.GOLD / "tables_small.npz")
    assert len(cases) >= 40
    for c in cases:
        if c.get("raises"):
            with pytest.raises(OverflowError):
                OC.build_bit_table(c["weights"], c["max_mass"], c["compression"])
            continue
        want = arrays[c["key"]]
        got = OC.build_bit_table(c["weights"], c["max_mass"], c["compression"])
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(got, want), c
        assert Hh.sha(got) == c["sha256"]
        if min(c["weights"][1:]) >= c["compression"]:
            assert np.array_equal(OP.build_bit_table_closed_form(c["weights"], c["max_mass"], c["compression"]), want), c


def test_byte_tables_match_reference():
    """set_up_mass_table (mass_table.py:292-316): the numpy restatement against reference-made tables."""
    cases = Hh.load_json("mass_tables_small.json")
    arrays = np.load(Hh.GOLD / "mass_tables_small.npz")
    assert len(cases) >= 10
    for c in cases:
        got = OP.build_mass_table(c["weights"], c["max_mass"])
        assert got.dtype == np.uint8 and np.array_equal(got, arrays[c["key"]]), c
        assert Hh.sha(got) == c["sha256"]


@pytest.mark.parametrize("name", ["acgu", "quirk_365045", "full"])
def test_big_table_sha(name):
    doc = Hh.load_json("tables_sha.json")[name]
    t = OC.build_bit_table(doc["weights"], doc["max_mass"], doc["compression"])
    assert list(t.shape) == doc["shape"]
    assert Hh.sha(t) == doc["sha256"]
    if name == "quirk_365045":
        assert doc["last_word_all_zero"] and (t[:, -1] == 0).all()
    if name == "full":
        assert doc["sha256"] == "fbbef632442646b71cf69f6a2eca185996c9d34b59f4ea95401ce0ed9a1bc266"


def test_explain_small_python_and_c():
    cases = Hh.load_json("explain_small.json.gz")
    tables = {}
    differ = 0
    for c in cases:
        w = c["weights"]
        tab = tables.setdefault(tuple(w), OC.build_bit_table(w, max(w) * 35, 32))
        rows = [OP.Row(m, im, rt) for m, im, rt in zip(w, c["is_mod"], c["rates"])]
        ind = OP.individual_budgets(rows, c["max_len"])
        target, thr = OP.integerise(c["mass"], c["threshold"], 1e-3, c["tolerance"])
        mm = math.inf if c["max_modifications"] is None else c["max_modifications"]
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            gold = c[f"solutions_{tag}"]
            if gold == "NotImplementedError":
                with pytest.raises(NotImplementedError):
                    OP.explain_solutions(c["mass"], tab, rows, c["max_len"], 32, 1e-3, c["tolerance"], mm, c["threshold"], memo)
                with pytest.raises(NotImplementedError):
                    OC.explain(tab, 32, w, c["is_mod"], ind, target, thr, c["max_modifications"], memo)
                continue
            got = OP.explain_solutions(c["mass"], tab, rows, c["max_len"], 32, 1e-3, c["tolerance"], mm, c["threshold"], memo)
            assert got == gold, (c, tag)
            r, off, _ = OC.explain(tab, 32, w, c["is_mod"], ind, target, thr, c["max_modifications"], memo)
            assert [[w[x] for x in r[off[i]:off[i + 1]]] for i in range(len(off) - 1)] == gold, (c, tag)
            names = {m: [f"n{m}"] for m in w[1:]}
            named = OP.solutions_to_names(got, names)
            want_named = c[f"named_{tag}"]
            assert (None if named is None else sorted(list(t) for t in named)) == want_named
        differ += c["solutions_memo"] != c["solutions_nomemo"]
        if c["is_valid"] == "NotImplementedError":
            with pytest.raises(NotImplementedError):
                OP.is_valid_mass(c["mass"], tab, 32, 1e-3, c["tolerance"], c["threshold"])
            with pytest.raises(NotImplementedError):
                OC.is_valid(tab, 32, target, thr)
        else:
            assert OP.is_valid_mass(c["mass"], tab, 32, 1e-3, c["tolerance"], c["threshold"]) == c["is_valid"]
            assert OC.is_valid(tab, 32, target, thr) == c["is_valid"]
        if "bound_lower" in c:
            for d in ("lower", "upper"):
                want = c[f"bound_{d}"]
                mass = c["mass"]
                args = (tab, rows, c["max_len"], mass, c["bound_obs_mass"], 0.5, 32, 1e-3, c["tolerance"], d)
                t2 = int(round(mass / 1e-3, 0))
                thr2 = int(np.ceil(c["tolerance"] * c["bound_obs_mass"] / 1e-3))
                if want == "NotImplementedError":
                    with pytest.raises(NotImplementedError):
                        OP.sequence_length_bound(*args)
                    continue
                assert OP.sequence_length_bound(*args) == want, (c, d)
                assert OC.length_bound(tab, 32, w, c["is_mod"], ind, t2, thr2, round(0.5 * c["max_len"]), c["max_len"], d) == want
    assert differ > 100  # the first-visit memo quirk is really exercised


@pytest.fixture(scope="module")
def full_table():
    w = Hh.full_weights()
    return w, OC.build_bit_table(w, max(w) * 35, 32)


def _full_rows(universal_rate=0.5):
    df = M.EXPLANATION_MASSES
    ims = df.get_column("tolerated_integer_masses").to_list()
    rates = dict(zip(ims, df.get_column("modification_rate").to_list()))
    rows = [OP.Row(0, False, 0.0)]
    for m in sorted(ims):
        mod = M._INT_MASS_IS_MOD[m]
        rows.append(OP.Row(m, mod, min(rates[m], universal_rate) if mod else rates[m]))
    return rows


def test_explain_full_alphabet_unit_cases(full_table):
    w, tab = full_table
    gold = Hh.load_json("explain_full.json")
    assert gold["table_sha256"] == Hh.sha(tab)
    rows = _full_rows()
    is_mod = [r.is_modification for r in rows]
    for c in gold["unit_test_cases"]:
        ind = OP.individual_budgets(rows, c["max_len"])
        target, thr = OP.integerise(c["mass"], None, M.TOLERANCE, c["tolerance"])
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            r, off, _ = OC.explain(tab, 32, w, is_mod, ind, target, thr, c["max_modifications"], memo)
            sols = [[w[x] for x in r[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
            named = OP.solutions_to_names(sols, M._INT_MASS_NAMES)
            assert (None if named is None else len(named)) == c[f"n_{tag}"], (c["seq"], c["tolerance"], tag)
            assert Hh.digest(named) == c[f"digest_{tag}"]
            assert tuple(c["seq"]) in named
            if f"set_{tag}" in c:
                assert sorted(list(t) for t in named) == c[f"set_{tag}"]
        if len(c["seq"]) <= 4:  # the pure-Python port too, where it is quick
            sols = OP.explain_solutions(c["mass"], tab, rows, c["max_len"], 32, M.TOLERANCE, c["tolerance"], c["max_modifications"], None, True)
            assert Hh.digest(OP.solutions_to_names(sols, M._INT_MASS_NAMES)) == c["digest_memo"]
            rec = OP.explain_mass_with_recursion(c["mass"], rows, c["max_len"], M.TOLERANCE, c["tolerance"], M._INT_MASS_IS_MOD, c["max_modifications"])
            assert Hh.digest(OP.solutions_to_names(rec, M._INT_MASS_NAMES)) == c["digest_recursion"]


def test_explain_full_alphabet_random_and_validity(full_table):
    w, tab = full_table
    gold = Hh.load_json("explain_full.json")
    rows = _full_rows()
    is_mod = [r.is_modification for r in rows]
    for c in gold["random_cases"]:
        ind = OP.individual_budgets(rows, c["max_len"])
        target, thr = OP.integerise(c["mass"], c["threshold"], M.TOLERANCE, 10e-6)
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            r, off, _ = OC.explain(tab, 32, w, is_mod, ind, target, thr, c["max_modifications"], memo)
            sols = [[w[x] for x in r[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
            assert Hh.digest(OP.solutions_to_names(sols, M._INT_MASS_NAMES)) == c[f"digest_{tag}"], c
    for c in gold["validity_cases"]:
        target, thr = OP.integerise(c["mass"], c["threshold"], M.TOLERANCE, 10e-6)
        if c["valid"] == "NotImplementedError":
            with pytest.raises(NotImplementedError):
                OC.is_valid(tab, 32, target, thr)
        else:
            assert OC.is_valid(tab, 32, target, thr) == c["valid"], c
