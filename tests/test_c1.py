"""BASELINE config C1 (the reference's tests/testcases/test_02 fixture) and rows N3 / N4 against reference-made goldens.

tests/golden/c1.json was produced by oracle/gen_golden_c1.py from the reference's UNMODIFIED function bodies
(classify_fragments, Predictor.filter_by_explanation and everything below it, the table's alphabet reduction) on
the fixture copied under tests/golden/c1/.  CPU tests pin the oracle and the host-side pair generator to it; GPU
tests compare the CUDA path — batched entries, scalar entries, and the reference's own prediction.py driving the
drop-in modules — with the same file.
"""
import pathlib
import sys
import types

import numpy as np
import pytest
import yaml

from helpers import load_json
from oracle import oracle_c as OC
from oracle import oracle_py as OP
from spectrseqtools_b200 import _frame as pl
from spectrseqtools_b200 import alphabet_reduction as AR
from spectrseqtools_b200 import masses as M

FIX = pathlib.Path(__file__).resolve().parent / "golden" / "c1"
CASES = {c["name"]: c for c in load_json("c1.json")}


def _names_of(weights):
    return [n for w in weights if w for n in M._INT_MASS_NAMES[w]]


def _frame_for(names):
    col = M.EXPLANATION_MASSES.get_column("nucleoside").to_list()
    df = M.EXPLANATION_MASSES.filter([n in names for n in col])
    rates = [1.0 if n in M.UNMODIFIED_BASES else r for n, r in zip(df.get_column("nucleoside").to_list(), df.get_column("modification_rate").to_list())]
    return df.with_columns(pl.Series("modification_rate", rates))


def _fragments():
    fr = pl.read_csv(FIX / "fragments.tsv", separator="\t")
    return fr.with_columns(pl.col("true_mass_with_backbone").alias("true_mass"))


def _named(sol_rows, weights):
    """oracle rows -> sorted list of name-sorted lists (what Explanation stores)."""
    return sorted(sorted(M._INT_MASS_NAMES[weights[r]][0] for r in rec) for rec in sol_rows)


# ---------------------------------------------------------------- CPU: the oracle and the host pair generator
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_oracle_reproduces_every_c1_call(name):
    case = CASES[name]
    tol = case["tolerance"]
    max_len = case["seq"]["max_len"]
    mm = round(case["seq"]["modification_rate"] * max_len)
    tabs = {}
    for n_rows, (weights, rates) in case["table_states"].items():
        tab = OC.build_bit_table(weights, max(weights) * 35, 32)
        rows = [OP.Row(w, bool(w and M._INT_MASS_IS_MOD[w]), r) for w, r in zip(weights, rates)]
        tabs[int(n_rows)] = (weights, tab, rows, OP.individual_budgets(rows, max_len))
    for mass, thr, want, n_rows in case["classify_calls"] + case["revalidate_calls"]:
        weights, tab, _rows, _ind = tabs[n_rows]
        assert OP.is_valid_mass(mass, tab, 32, 1e-3, tol, thr) == want
    checked = 0
    for diff, thr, want, n_rows in case["explain_calls"]:
        weights, tab, rows, ind = tabs[n_rows]
        t, h = OP.integerise(diff, thr, 1e-3, tol)
        r, off, zero = OC.explain(tab, 32, weights, [x.is_modification for x in rows], ind, t, h, mm, True)
        sols = [tuple(int(x) for x in r[off[i]:off[i + 1]]) for i in range(len(off) - 1)]
        got = None if not sols else _named([s for s in sols if s], weights)
        assert got == want, (diff, thr)
        checked += 1
    assert checked == len(case["explain_calls"]) >= 400


@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_first_round_calls_are_the_reference_window(name):
    """N3: the ladder pairs + singletons of the classified frame are exactly the reference's first-round calls
    (same differences and l1 thresholds, bit for bit, in the same order)."""
    case = CASES[name]
    cl = case["classified"]
    su, obs, brk, single = cl["standard_unit_mass"], cl["observed_mass"], cl["breakage"], cl["is_singleton"]
    first_round = [c for c in case["explain_calls"] if c[3] == len(case["start_weights"])]
    max_weight = AR._max_weight(_frame_for(set(case["alphabet"])))
    calls = []
    for tag in ("START", "END"):
        idx = [i for i, b in enumerate(brk) if tag in b]
        s_su, s_obs = [su[i] for i in idx], [obs[i] for i in idx]
        for s, e in AR.ladder_pairs(s_su, max_weight):
            calls.append((s_su[e] - s_su[s], case["tolerance"] * (s_obs[s] + s_obs[e])))
    calls += [(su[i], case["tolerance"] * obs[i]) for i in range(len(su)) if single[i]]
    assert calls == [(c[0], c[1]) for c in first_round]


# ---------------------------------------------------------------- GPU
def _dp(case, names=None):
    from spectrseqtools_b200 import mass_table as MT

    s = case["seq"]
    seq = MT.SequenceInformation(max_len=s["max_len"], su_mass=s["su_mass"], obs_mass=s["obs_mass"], modification_rate=s["modification_rate"])
    return MT.DynamicProgrammingTable(_frame_for(set(names or case["alphabet"])), 32, case["tolerance"], 1e-3, seq)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_c1_classification_matches_the_reference(name):
    from spectrseqtools_b200 import fragment_classification as FC

    case = CASES[name]
    dp = _dp(case)
    assert [m.mass for m in dp.masses] == case["start_weights"]
    breakage = {int(k): v for k, v in case["breakage"].items()}
    got = FC.classify_fragments(_fragments(), dp, breakage, intensity_cutoff=M.DEFAULT_INTENSITY_CUTOFF)
    for colname, want in case["classified"].items():
        assert got.get_column(colname).to_list() == want, colname
    # every (fragment x breakage) validity answer, in the reference's call order (breakage-major)
    res = FC.classify_observed(_fragments().get_column("observed_mass").to_list(), dp, breakage)
    assert [bool(x) for x in res.valid.reshape(-1)] == [c[2] for c in case["classify_calls"]]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_c1_every_explanation_call_matches_the_reference(name):
    from spectrseqtools_b200 import common, mass_explanation as ME

    case = CASES[name]
    for n_rows, (weights, _rates) in case["table_states"].items():
        dp = _dp(case, _names_of(weights))
        assert [m.mass for m in dp.masses] == weights
        calls = [c for c in case["explain_calls"] if c[3] == int(n_rows)]
        got = common.calculate_explanations_batch([c[0] for c in calls], [c[1] for c in calls], dp)
        for c, g in zip(calls, got):
            assert (None if g is None else sorted(list(e.nucleosides) for e in g)) == c[2], c[:2]
        for c in calls[:: max(1, len(calls) // 25)]:  # the scalar, reference-shaped entry on a subsample
            g = common.calculate_explanations(c[0], c[1], dp)
            assert (None if g is None else sorted(list(e.nucleosides) for e in g)) == c[2], c[:2]
        reval = [c for c in case["revalidate_calls"] if c[3] == int(n_rows)]
        if reval:
            codes = ME.are_valid_masses([c[0] for c in reval], dp, [c[1] for c in reval])
            assert [bool(x == 1) for x in codes] == [c[2] for c in reval]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_c1_fixed_point_matches_the_reference(name):
    """N4: explanation -> alphabet reduction -> table rebuild -> re-validation, to the reference's fixed point."""
    case = CASES[name]
    dp = _dp(case)
    cl = case["classified"]
    alive, expl = AR.filter_by_explanation(cl["standard_unit_mass"], cl["observed_mass"], cl["breakage"], cl["is_singleton"], dp,
                                           _frame_for(set(case["alphabet"])))
    assert [int(i) for i in alive] == case["final_orig_index"]
    assert [m.mass for m in dp.masses] == case["final_weights"]
    assert [list(m.names) for m in dp.masses] == case["final_names"]
    want = {k: v for k, v in case["explanations"]}
    assert set(expl) == set(want)
    for k, v in expl.items():
        assert (None if v is None else sorted(list(e.nucleosides) for e in v)) == want[k], k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_c1_device_ladder_first_round_is_the_reference_window(name):
    """N3 on the device: the generated calls (differences, l1 thresholds; START pairs, END pairs, singletons) are the
    reference's first-round calls bit for bit and in order, and the surviving dict entries are the reference's."""
    case = CASES[name]
    dp = _dp(case)
    cl = case["classified"]
    lad = AR.DeviceLadder(cl["standard_unit_mass"], cl["observed_mass"], cl["breakage"], cl["is_singleton"], dp, _frame_for(set(case["alphabet"])))
    names = lad.round()
    keys, thr, fl = lad.calls()
    first_round = [c for c in case["explain_calls"] if c[3] == len(case["start_weights"])]
    assert [float(k) for k in keys] == [c[0] for c in first_round]
    assert [float(t) for t in thr] == [c[1] for c in first_round]
    # the host generator + batched enumeration build the same dict (keys, order of first insertion, surviving values)
    want = AR.collect_diff_explanations(cl["standard_unit_mass"], cl["observed_mass"], cl["breakage"], cl["is_singleton"], dp, _frame_for(set(case["alphabet"])))
    lad.round()
    got = lad.explanations()
    assert list(got) == list(want)
    for k in want:
        assert (None if got[k] is None else sorted(e.nucleosides for e in got[k])) == (None if want[k] is None else sorted(e.nucleosides for e in want[k])), k
    assert names == AR.observed_nucleotides(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_c1_device_fixed_point_matches_the_reference(name):
    """N4 with the frame resident on the device: the reference's fixed point, surviving fragments and final explanations."""
    case = CASES[name]
    dp = _dp(case)
    cl = case["classified"]
    alive, expl = AR.filter_by_explanation_device(cl["standard_unit_mass"], cl["observed_mass"], cl["breakage"], cl["is_singleton"], dp,
                                                  _frame_for(set(case["alphabet"])))
    assert [int(i) for i in alive] == case["final_orig_index"]
    assert [m.mass for m in dp.masses] == case["final_weights"]
    assert [list(m.names) for m in dp.masses] == case["final_names"]
    want = {k: v for k, v in case["explanations"]}
    assert set(expl) == set(want)
    for k, v in expl.items():
        assert (None if v is None else sorted(list(e.nucleosides) for e in v)) == want[k], k


@pytest.mark.gpu
def test_device_ladder_random_frames_equal_the_host_generator():
    """Dense random ladders with repeated masses (equal keys on both sides and among the singletons), fragments that die
    in a round, empty sides: pairs, thresholds, dict and row mask equal the host generator's, round after round."""
    rng = np.random.default_rng(5)
    case = CASES["mods12"]
    frame = _frame_for(set(case["alphabet"]))
    w = [x for x in case["start_weights"] if x]
    for trial in range(6):
        dp = _dp(case)
        n = [0, 1, 2, 40, 300, 1500][trial]
        ladders = [np.cumsum(rng.choice(w, size=25)) * 1e-3 for _ in range(n // 25 + 1)]  # (25 nucleotides stay inside the table)
        su = np.concatenate(ladders)[:n]
        su = np.sort(np.concatenate([su, su[: n // 3]]))  # a third of the masses twice
        obs = su + rng.choice([0.0, 18.0105, 97.9769], size=len(su))
        brk = [["START_c/y", "c/y_END", "START_END", "c/y_c/y"][k] for k in rng.integers(0, 4, size=len(su))]
        single = (rng.random(len(su)) < 0.5) & (np.arange(len(su)) < 12)  # singletons are light fragments (heavy ones leave the table)
        lad = AR.DeviceLadder(su, obs, brk, single, dp, frame)
        names = lad.round()
        want = AR.collect_diff_explanations(su, obs, brk, single, dp, frame)
        got = lad.explanations()
        assert list(got) == list(want), trial
        for k in want:
            assert (None if got[k] is None else sorted(e.nucleosides for e in got[k])) == (None if want[k] is None else sorted(e.nucleosides for e in want[k])), (trial, k)
        assert names == AR.observed_nucleotides(want), trial
        if len(su):
            keep = AR.reduce_alphabet(names, su, obs, dp)  # host path: rebuilds the table, validity batch
            assert lad.revalidate() == int(keep.sum())
            assert np.array_equal(lad.alive(), keep), trial
            want2 = AR.collect_diff_explanations(su[keep], obs[keep], [b for b, k in zip(brk, keep) if k], single[keep], dp, frame)
            lad.round()
            got2 = lad.explanations()
            assert list(got2) == list(want2), trial


@pytest.mark.gpu
def test_device_ladder_raises_beyond_the_table_like_upstream():
    """A singleton whose window reaches beyond the table: calculate_explanations raises NotImplementedError upstream
    (mass_explanation.py:134-138) in the middle of the round; the device round and the host generator raise it too."""
    case = CASES["mods12"]
    frame = _frame_for(set(case["alphabet"]))
    dp = _dp(case)
    su = np.array([329.0525, 658.105, 40000.0])
    obs = su + 18.0105
    brk = ["START_c/y", "START_c/y", "START_c/y"]
    single = np.array([False, False, True])
    with pytest.raises(NotImplementedError):
        AR.collect_diff_explanations(su, obs, brk, single, dp, frame)
    lad = AR.DeviceLadder(su, obs, brk, single, dp, frame)
    with pytest.raises(NotImplementedError):
        lad.round()
    # without the heavy singleton the same frame goes through
    lad = AR.DeviceLadder(su[:2], obs[:2], brk[:2], single[:2], dp, frame)
    assert lad.round() == AR.observed_nucleotides(AR.collect_diff_explanations(su[:2], obs[:2], brk[:2], single[:2], dp, frame))


def _reference_checkout():
    for cand in (pathlib.Path("/root/reference"), pathlib.Path(__file__).resolve().parents[1] / "baseline" / "_ref"):
        if (cand / "spectrseqtools" / "prediction.py").is_file():
            return cand
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["acgu", "mods12"])
def test_reference_prediction_module_runs_unchanged_on_the_drop_in(name):
    """The reference's OWN prediction.py (imported as spectrseqtools.prediction through the alias package, unmodified)
    drives the CUDA path: Predictor.filter_by_explanation on the C1 fixture gives the reference's own answer."""
    if _reference_checkout() is None:
        pytest.skip("no reference checkout (baseline/_ref is made by __graft_entry__.build() in the build container)")
    if "pulp" not in sys.modules:
        try:
            import pulp  # noqa: F401
        except ImportError:  # the MILP layer is out of scope: its import only has to resolve
            stub = types.ModuleType("pulp")
            for n in "LpProblem LpMinimize LpInteger LpContinuous LpVariable lpSum getSolver".split():
                setattr(stub, n, None)
            sys.modules["pulp"] = stub
    import spectrseqtools
    from spectrseqtools.prediction import Predictor
    import spectrseqtools.prediction as upstream

    assert "b200" not in upstream.__file__ and upstream.__file__.endswith("prediction.py")
    assert spectrseqtools.mass_explanation.__name__ == "spectrseqtools_b200.mass_explanation"
    from spectrseqtools.fragment_classification import classify_fragments
    import polars  # the stand-in (or the real one): whatever the alphabet frame is made of

    case = CASES[name]
    dp = _dp(case)
    breakage = {int(k): v for k, v in case["breakage"].items()}
    frags = classify_fragments(_fragments(), dp, breakage, intensity_cutoff=M.DEFAULT_INTENSITY_CUTOFF)
    frags = frags.with_row_index(name="orig_index").sort("standard_unit_mass").with_row_index(name="index")  # prediction.py:68-72
    final, expl = Predictor(dp_table=dp, explanation_masses=_frame_for(set(case["alphabet"]))).filter_by_explanation(frags)
    assert final.get_column("orig_index").to_list() == case["final_orig_index"]
    assert [m.mass for m in dp.masses] == case["final_weights"]
    want = {k: v for k, v in case["explanations"]}
    assert set(expl) == set(want)
    for k, v in expl.items():
        assert (None if v is None else sorted(list(e.nucleosides) for e in v)) == want[k], k
